// Host-side binary32 vector / matrix types of the front end.
//
// Mirrors the reference's `Vector` / `Matrix<3>` API (include/tracer/Vector.h:5-30, include/tracer/Matrix.h:6-30)
// so code written against the reference reads the same, but it is a fresh implementation: plain structs, no bounds-
// checked accessor (the reference's Vector::operator[] check is 33 % of its CPU time).
// The expression ORDER of every operation follows SURVEY.md App. A-1 exactly, and this directory is compiled with
// -ffp-contract=off and without -march=native, so results are bit-identical to the reference's x86-64 build.
#pragma once
#include <cmath>
#include <cstddef>
#include <initializer_list>

namespace crt {

struct Vector {
  float x = 0.0f, y = 0.0f, z = 0.0f;
  Vector() = default;
  Vector(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
  float &operator[](unsigned i) { return i == 0 ? x : (i == 1 ? y : z); }
  const float &operator[](unsigned i) const { return i == 0 ? x : (i == 1 ? y : z); }
  Vector operator-(const Vector &o) const { return {x - o.x, y - o.y, z - o.z}; }
  Vector operator+(const Vector &o) const { return {x + o.x, y + o.y, z + o.z}; }
  Vector &operator+=(const Vector &o) {
    x += o.x;
    y += o.y;
    z += o.z;
    return *this;
  }
  // (a0*b0 + a1*b1) + a2*b2   -- Vector.cpp:57-59
  float dot(const Vector &o) const { return x * o.x + y * o.y + z * o.z; }
  // cross product, spelled operator* in the reference -- Vector.cpp:61-65
  Vector operator*(const Vector &o) const { return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x}; }
  Vector operator*(float s) const { return {x * s, y * s, z * s}; }
  friend Vector operator*(float s, const Vector &v) { return {s * v.x, s * v.y, s * v.z}; }
  bool operator==(const Vector &o) const { return x == o.x && y == o.y && z == o.z; }
  float length() const { return std::sqrt(x * x + y * y + z * z); }  // Vector.cpp:114-117
  void normalize() {                                                  // Vector.cpp:97-106
    float len = length();
    if (len == 0) return;
    len = 1.0f / len;
    x *= len;
    y *= len;
    z *= len;
  }
  Vector getNormalized() const {
    Vector t(*this);
    t.normalize();
    return t;
  }
  Vector reflect(const Vector &n) const { return *this - (2 * this->dot(n)) * n; }  // Vector.cpp:119-122
};
using Color = Vector;
using Albedo = Vector;

// Row-major 3x3, row-vector convention (v * M), Matrix.h:137-142.
struct Matrix3 {
  float m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  Matrix3() = default;
  Matrix3(std::initializer_list<float> v) {
    size_t k = 0;
    for (float f : v) {
      if (k >= 9) break;
      m[k / 3][k % 3] = f;
      k++;
    }
  }
  static Matrix3 identity() { return Matrix3{1, 0, 0, 0, 1, 0, 0, 0, 1}; }
  float *operator[](unsigned r) { return m[r]; }
  const float *operator[](unsigned r) const { return m[r]; }
};
static const Matrix3 IDENTITY_MATRIX = Matrix3::identity();

inline Vector operator*(const Vector &l, const Matrix3 &r) {
  return {l.x * r.m[0][0] + l.y * r.m[1][0] + l.z * r.m[2][0],  //
          l.x * r.m[0][1] + l.y * r.m[1][1] + l.z * r.m[2][1],  //
          l.x * r.m[0][2] + l.y * r.m[1][2] + l.z * r.m[2][2]};
}
// left = left * right with the accumulation order of Matrix.h:144-157 (0 + a*b + a*b + a*b)
inline Matrix3 &operator*=(Matrix3 &l, const Matrix3 &r) {
  Matrix3 out;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      float acc = 0;
      for (int k = 0; k < 3; k++) acc += l.m[i][k] * r.m[k][j];
      out.m[i][j] = acc;
    }
  l = out;
  return l;
}

}  // namespace crt
