// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.
// PARITY UNPINNED: pybox2d/Box2D is not present in /root/reference nor installable here; this is a
// from-memory restatement of Box2D 2.3.x (box2d-py, un-pinned at /root/reference/setup.py:5), the
// library gym_kilobots calls at envs/kilobots_env.py:187.  See DESIGN.md "Oracle".
//
// kbo_math.h -- float32 vector math restating Box2D's b2Math.h (b2Vec2, b2Rot, b2Transform, b2Sweep,
// b2Mat22) operation by operation, so that expression order (and therefore rounding) matches.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

namespace kbo {

// b2Settings.h
constexpr float kPi = 3.14159265359f;
constexpr float kEpsilon = FLT_EPSILON;
constexpr float kMaxFloat = FLT_MAX;
constexpr float kLinearSlop = 0.005f;
constexpr float kAngularSlop = 2.0f / 180.0f * kPi;
constexpr float kPolygonRadius = 2.0f * kLinearSlop;
constexpr float kAabbExtension = 0.1f;
constexpr float kAabbMultiplier = 2.0f;
constexpr float kVelocityThreshold = 1.0f;
constexpr float kMaxLinearCorrection = 0.2f;
constexpr float kMaxAngularCorrection = 8.0f / 180.0f * kPi;
constexpr float kMaxTranslation = 2.0f;
constexpr float kMaxTranslationSquared = kMaxTranslation * kMaxTranslation;
constexpr float kMaxRotation = 0.5f * kPi;
constexpr float kMaxRotationSquared = kMaxRotation * kMaxRotation;
constexpr float kBaumgarte = 0.2f;
constexpr float kToiBaumgarte = 0.75f;
constexpr float kTimeToSleep = 0.5f;
constexpr float kLinearSleepTolerance = 0.01f;
constexpr float kAngularSleepTolerance = 2.0f / 180.0f * kPi;
constexpr int kMaxSubSteps = 8;
constexpr int kMaxTOIContacts = 32;
constexpr int kMaxManifoldPoints = 2;
constexpr int kMaxPolygonVertices = 8;

// --- trig ------------------------------------------------------------------------------------
// Box2D's b2Rot::Set calls libm sinf/cosf.  CUDA's sinf/cosf are not bit-identical to glibc's, so
// the oracle and the kernel both evaluate the SAME double-precision algorithm (Cody-Waite
// reduction by pi/2 + fdlibm-style kernels, plain IEEE mul/add, no FMA) and round to float.
// tests/test_oracle_trig.py measures its agreement with glibc sinf/cosf.  Define KBO_LIBM_TRIG to
// use libm instead (sensitivity studies only; breaks bit-parity with the kernel).
inline void SinCosD(double x, double* s, double* c) {
  const double kd = rint(x * 6.36619772367581382433e-01);
  const double r = (x - kd * 1.57079632673412561417e+00) - kd * 6.07710050650619224932e-11;
  const double z = r * r;
  const double ps = -1.66666666666666324348e-01 +
                    z * (8.33333333332248946124e-03 +
                         z * (-1.98412698298579493134e-04 +
                              z * (2.75573137070700676789e-06 +
                                   z * (-2.50507602534068634195e-08 + z * 1.58969099521155010221e-10))));
  const double sn = r + (r * z) * ps;
  const double pc = 4.16666666666666019037e-02 +
                    z * (-1.38888888888741095749e-03 +
                         z * (2.48015872894767294178e-05 +
                              z * (-2.75573143513906633035e-07 +
                                   z * (2.08757232129817482790e-09 + z * -1.13596475577881948265e-11))));
  const double cs = (1.0 - 0.5 * z) + (z * z) * pc;
  const long long k = (long long)kd;
  switch ((int)(k & 3)) {
    case 0: *s = sn; *c = cs; break;
    case 1: *s = cs; *c = -sn; break;
    case 2: *s = -sn; *c = -cs; break;
    default: *s = -cs; *c = sn; break;
  }
}

inline void SinCos(float a, float* s, float* c) {
#ifdef KBO_LIBM_TRIG
  *s = sinf(a);
  *c = cosf(a);
#else
  double sd, cd;
  SinCosD((double)a, &sd, &cd);
  *s = (float)sd;
  *c = (float)cd;
#endif
}

// --- b2Math.h ---------------------------------------------------------------------------------
struct Vec2 {
  float x, y;
  Vec2() : x(0.0f), y(0.0f) {}
  Vec2(float x_, float y_) : x(x_), y(y_) {}
  void Set(float x_, float y_) { x = x_; y = y_; }
  void SetZero() { x = 0.0f; y = 0.0f; }
  Vec2 operator-() const { return Vec2(-x, -y); }
  void operator+=(const Vec2& v) { x += v.x; y += v.y; }
  void operator-=(const Vec2& v) { x -= v.x; y -= v.y; }
  void operator*=(float a) { x *= a; y *= a; }
  float Length() const { return sqrtf(x * x + y * y); }
  float LengthSquared() const { return x * x + y * y; }
  float Normalize() {
    float length = Length();
    if (length < kEpsilon) return 0.0f;
    float invLength = 1.0f / length;
    x *= invLength;
    y *= invLength;
    return length;
  }
};
inline Vec2 operator+(const Vec2& a, const Vec2& b) { return Vec2(a.x + b.x, a.y + b.y); }
inline Vec2 operator-(const Vec2& a, const Vec2& b) { return Vec2(a.x - b.x, a.y - b.y); }
inline Vec2 operator*(float s, const Vec2& a) { return Vec2(s * a.x, s * a.y); }
inline float Dot(const Vec2& a, const Vec2& b) { return a.x * b.x + a.y * b.y; }
inline float Cross(const Vec2& a, const Vec2& b) { return a.x * b.y - a.y * b.x; }
inline Vec2 Cross(const Vec2& a, float s) { return Vec2(s * a.y, -s * a.x); }
inline Vec2 Cross(float s, const Vec2& a) { return Vec2(-s * a.y, s * a.x); }
inline float DistanceSquared(const Vec2& a, const Vec2& b) {
  Vec2 c = a - b;
  return Dot(c, c);
}
inline float Min(float a, float b) { return a < b ? a : b; }
inline float Max(float a, float b) { return a > b ? a : b; }
inline int Min(int a, int b) { return a < b ? a : b; }
inline int Max(int a, int b) { return a > b ? a : b; }
inline Vec2 Min(const Vec2& a, const Vec2& b) { return Vec2(Min(a.x, b.x), Min(a.y, b.y)); }
inline Vec2 Max(const Vec2& a, const Vec2& b) { return Vec2(Max(a.x, b.x), Max(a.y, b.y)); }
inline float Clamp(float a, float lo, float hi) { return Max(lo, Min(a, hi)); }
inline float Abs(float a) { return a > 0.0f ? a : -a; }

struct Mat22 {
  Vec2 ex, ey;
  void SetZero() { ex.SetZero(); ey.SetZero(); }
  Mat22 GetInverse() const {
    float a = ex.x, b = ey.x, c = ex.y, d = ey.y;
    Mat22 B;
    float det = a * d - b * c;
    if (det != 0.0f) det = 1.0f / det;
    B.ex.x = det * d;
    B.ey.x = -det * b;
    B.ex.y = -det * c;
    B.ey.y = det * a;
    return B;
  }
};
inline Vec2 Mul(const Mat22& A, const Vec2& v) {
  return Vec2(A.ex.x * v.x + A.ey.x * v.y, A.ex.y * v.x + A.ey.y * v.y);
}

struct Rot {
  float s, c;
  Rot() : s(0.0f), c(1.0f) {}
  explicit Rot(float angle) { Set(angle); }
  void Set(float angle) { SinCos(angle, &s, &c); }
};
inline Rot MulT(const Rot& q, const Rot& r) {
  Rot qr;
  qr.s = q.c * r.s - q.s * r.c;
  qr.c = q.c * r.c + q.s * r.s;
  return qr;
}
inline Vec2 Mul(const Rot& q, const Vec2& v) { return Vec2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
inline Vec2 MulT(const Rot& q, const Vec2& v) { return Vec2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }

struct Xf {
  Vec2 p;
  Rot q;
};
inline Vec2 Mul(const Xf& T, const Vec2& v) {
  float x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
  float y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
  return Vec2(x, y);
}
inline Vec2 MulT(const Xf& T, const Vec2& v) {
  float px = v.x - T.p.x;
  float py = v.y - T.p.y;
  float x = (T.q.c * px + T.q.s * py);
  float y = (-T.q.s * px + T.q.c * py);
  return Vec2(x, y);
}
inline Xf MulT(const Xf& A, const Xf& B) {
  Xf C;
  C.q = MulT(A.q, B.q);
  C.p = MulT(A.q, B.p - A.p);
  return C;
}

struct Sweep {
  Vec2 localCenter;
  Vec2 c0, c;
  float a0 = 0.0f, a = 0.0f;
  float alpha0 = 0.0f;
  void GetTransform(Xf* xf, float beta) const {
    xf->p = (1.0f - beta) * c0 + beta * c;
    float angle = (1.0f - beta) * a0 + beta * a;
    xf->q.Set(angle);
    xf->p -= Mul(xf->q, localCenter);
  }
  void Advance(float alpha) {
    float beta = (alpha - alpha0) / (1.0f - alpha0);
    c0 += beta * (c - c0);
    a0 += beta * (a - a0);
    alpha0 = alpha;
  }
  void Normalize() {
    float twoPi = 2.0f * kPi;
    float d = twoPi * floorf(a0 / twoPi);
    a0 -= d;
    a -= d;
  }
};

struct AABB {
  Vec2 lowerBound, upperBound;
  void Combine(const AABB& a, const AABB& b) {
    lowerBound = Min(a.lowerBound, b.lowerBound);
    upperBound = Max(a.upperBound, b.upperBound);
  }
  bool Contains(const AABB& aabb) const {
    bool result = true;
    result = result && lowerBound.x <= aabb.lowerBound.x;
    result = result && lowerBound.y <= aabb.lowerBound.y;
    result = result && aabb.upperBound.x <= upperBound.x;
    result = result && aabb.upperBound.y <= upperBound.y;
    return result;
  }
};
inline bool TestOverlap(const AABB& a, const AABB& b) {
  Vec2 d1 = b.lowerBound - a.upperBound;
  Vec2 d2 = a.lowerBound - b.upperBound;
  if (d1.x > 0.0f || d1.y > 0.0f) return false;
  if (d2.x > 0.0f || d2.y > 0.0f) return false;
  return true;
}

}  // namespace kbo
