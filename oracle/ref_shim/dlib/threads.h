// TEST INFRASTRUCTURE ONLY -- enough of <dlib/threads.h> for OdometryPipeline.h to be a complete type.
#pragma once
namespace dlib {
class multithreaded_object { public: virtual ~multithreaded_object() {} protected: void start() {} void wait() {} };
class mutex {};
class auto_mutex { public: explicit auto_mutex(mutex&) {} void unlock() {} };
}
