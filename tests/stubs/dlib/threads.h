// declaration-only stand-in for <dlib/threads.h> (syntax check of the adapters only)
#pragma once
namespace dlib {
class multithreaded_object { public: virtual ~multithreaded_object() {} protected: void start(); void wait(); };
class mutex {}; class auto_mutex { public: explicit auto_mutex(mutex&) {} void unlock() {} };
}
