#pragma once
#include <cstddef>
namespace dlib { template <typename T> class pipe { public: explicit pipe(size_t) {} bool enqueue(T&); bool dequeue(T&); }; }
