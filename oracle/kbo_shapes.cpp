// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see kbo_math.h).
// Restates b2PolygonShape::{SetAsBox,Set,ComputeAABB,ComputeMass}, b2CircleShape::{ComputeAABB,
// ComputeMass}, b2ChainShape::ComputeAABB of Box2D 2.3.x -- the shape code reached from
// gym_kilobots/lib/body.py:136-142 (box), :187-192 (circle), :245-251 (polygon) and
// envs/kilobots_env.py:46-51 (chain).
#include "kbo_world.h"

namespace kbo {

void Shape::SetAsBox(float hx, float hy) {
  type = kPolygon;
  radius = kPolygonRadius;
  count = 4;
  vertices[0].Set(-hx, -hy);
  vertices[1].Set(hx, -hy);
  vertices[2].Set(hx, hy);
  vertices[3].Set(-hx, hy);
  normals[0].Set(0.0f, -1.0f);
  normals[1].Set(1.0f, 0.0f);
  normals[2].Set(0.0f, 1.0f);
  normals[3].Set(-1.0f, 0.0f);
  centroid.SetZero();
}

static Vec2 ComputeCentroid(const Vec2* vs, int count) {
  Vec2 c(0.0f, 0.0f);
  float area = 0.0f;
  Vec2 pRef(0.0f, 0.0f);
  const float inv3 = 1.0f / 3.0f;
  for (int i = 0; i < count; ++i) {
    Vec2 p1 = pRef;
    Vec2 p2 = vs[i];
    Vec2 p3 = i + 1 < count ? vs[i + 1] : vs[0];
    Vec2 e1 = p2 - p1;
    Vec2 e2 = p3 - p1;
    float D = Cross(e1, e2);
    float triangleArea = 0.5f * D;
    area += triangleArea;
    c += (triangleArea * inv3) * ((p1 + p2) + p3);
  }
  c *= 1.0f / area;
  return c;
}

// b2PolygonShape::Set: weld, gift-wrap hull from the right-most point (CCW), normals, centroid.
void Shape::SetPolygon(const Vec2* verts, int n_in) {
  type = kPolygon;
  radius = kPolygonRadius;
  int n = Min(n_in, kMaxPolygonVertices);
  Vec2 ps[kMaxPolygonVertices];
  int tempCount = 0;
  for (int i = 0; i < n; ++i) {
    Vec2 v = verts[i];
    bool unique = true;
    for (int j = 0; j < tempCount; ++j) {
      if (DistanceSquared(v, ps[j]) < 0.5f * kLinearSlop) {  // sic: Box2D 2.3.0 compares d^2 with 0.5*slop
        unique = false;
        break;
      }
    }
    if (unique) ps[tempCount++] = v;
  }
  n = tempCount;
  if (n < 3) {
    SetAsBox(1.0f, 1.0f);
    return;
  }
  int i0 = 0;
  float x0 = ps[0].x;
  for (int i = 1; i < n; ++i) {
    float x = ps[i].x;
    if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) {
      i0 = i;
      x0 = x;
    }
  }
  int hull[kMaxPolygonVertices];
  int m = 0;
  int ih = i0;
  for (;;) {
    hull[m] = ih;
    int ie = 0;
    for (int j = 1; j < n; ++j) {
      if (ie == ih) {
        ie = j;
        continue;
      }
      Vec2 r = ps[ie] - ps[hull[m]];
      Vec2 v = ps[j] - ps[hull[m]];
      float c = Cross(r, v);
      if (c < 0.0f) ie = j;
      if (c == 0.0f && v.LengthSquared() > r.LengthSquared()) ie = j;
    }
    ++m;
    ih = ie;
    if (ie == i0) break;
  }
  count = m;
  for (int i = 0; i < m; ++i) vertices[i] = ps[hull[i]];
  for (int i = 0; i < m; ++i) {
    int i1 = i;
    int i2 = i + 1 < m ? i + 1 : 0;
    Vec2 edge = vertices[i2] - vertices[i1];
    normals[i] = Cross(edge, 1.0f);
    normals[i].Normalize();
  }
  centroid = ComputeCentroid(vertices, m);
}

void Shape::ComputeAABB(AABB* aabb, const Xf& xf) const {
  if (type == kCircle) {
    Vec2 c = xf.p + Mul(xf.q, p);
    aabb->lowerBound.Set(c.x - radius, c.y - radius);
    aabb->upperBound.Set(c.x + radius, c.y + radius);
  } else if (type == kPolygon) {
    Vec2 lower = Mul(xf, vertices[0]);
    Vec2 upper = lower;
    for (int i = 1; i < count; ++i) {
      Vec2 v = Mul(xf, vertices[i]);
      lower = Min(lower, v);
      upper = Max(upper, v);
    }
    Vec2 r(radius, radius);
    aabb->lowerBound = lower - r;
    aabb->upperBound = upper + r;
  } else {
    // b2ChainShape::ComputeAABB(childIndex): NO radius term (unlike b2EdgeShape::ComputeAABB).
    Vec2 a = Mul(xf, v1);
    Vec2 b = Mul(xf, v2);
    aabb->lowerBound = Min(a, b);
    aabb->upperBound = Max(a, b);
  }
}

void Shape::ComputeMass(MassData* md, float density) const {
  if (type == kCircle) {
    md->mass = density * kPi * radius * radius;
    md->center = p;
    md->I = md->mass * (0.5f * radius * radius + Dot(p, p));
    return;
  }
  if (type == kEdge) {
    md->mass = 0.0f;
    md->center.SetZero();
    md->I = 0.0f;
    return;
  }
  Vec2 center(0.0f, 0.0f);
  float area = 0.0f;
  float I = 0.0f;
  Vec2 s(0.0f, 0.0f);
  for (int i = 0; i < count; ++i) s += vertices[i];
  s *= 1.0f / count;
  const float k_inv3 = 1.0f / 3.0f;
  for (int i = 0; i < count; ++i) {
    Vec2 e1 = vertices[i] - s;
    Vec2 e2 = i + 1 < count ? vertices[i + 1] - s : vertices[0] - s;
    float D = Cross(e1, e2);
    float triangleArea = 0.5f * D;
    area += triangleArea;
    center += (triangleArea * k_inv3) * (e1 + e2);
    float ex1 = e1.x, ey1 = e1.y;
    float ex2 = e2.x, ey2 = e2.y;
    float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
    float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
    I += (0.25f * k_inv3 * D) * (intx2 + inty2);
  }
  md->mass = density * area;
  center *= 1.0f / area;
  md->center = center + s;
  md->I = density * I;
  md->I += md->mass * (Dot(md->center, md->center) - Dot(center, center));
}

}  // namespace kbo
