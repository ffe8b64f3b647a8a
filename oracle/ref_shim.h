// Test infrastructure only. Force-included (-include) when compiling the UNMODIFIED reference
// sources from /root/reference/SourceCode into oracle/_ref/: libstdc++ 13 does not declare
// std::sinf/cosf/sqrtf/powf, which the reference calls (RayTracer.cpp:342-345,396,406-409,
// Vector.cpp:132-134).  Nothing else is changed.
#pragma once
#include <cmath>
#include <math.h>
namespace std {
using ::cosf;
using ::powf;
using ::sinf;
using ::sqrtf;
}  // namespace std
