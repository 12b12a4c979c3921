// kb_narrow.cuh -- narrowphase, one lane per persistent contact.
//
// GPU restatement of the manifold routines Box2D 2.3.x runs inside b2Contact::Update (reached from
// b2World.Step, gym_kilobots/envs/kilobots_env.py:187): circle-circle, polygon-circle,
// polygon-polygon (2.3.1 brute-force max separation), chain-edge-circle, chain-edge-polygon
// (SURVEY.md Appendix B.5.1 / B.5.2).  Shapes live in the scene template (read-only, L1/L2
// resident); a lane keeps its manifold in registers and never materialises temporary polygons:
// transformed vertices are recomputed where needed, which is the same float32 arithmetic.
#pragma once
#include "kb_types.cuh"

namespace kb {

__device__ __forceinline__ V2 pvert(const ProxyConst* __restrict__ p, int i) { return mk(__ldg(&p->vx[i]), __ldg(&p->vy[i])); }
__device__ __forceinline__ V2 pnorm(const ProxyConst* __restrict__ p, int i) { return mk(__ldg(&p->nx[i]), __ldg(&p->ny[i])); }

__device__ __forceinline__ uint32_t make_id(uint32_t indexA, uint32_t indexB, uint32_t typeA, uint32_t typeB) {
  return (indexA & 0xFFu) | ((indexB & 0xFFu) << 8) | ((typeA & 0xFFu) << 16) | ((typeB & 0xFFu) << 24);
}
__device__ __forceinline__ uint32_t swap_id(uint32_t id) {
  // indexA<->indexB, typeA<->typeB
  return ((id >> 8) & 0xFFu) | ((id & 0xFFu) << 8) | ((id >> 8) & 0xFF0000u) | ((id & 0xFF0000u) << 8);
}

// circle centres are at the body origin for every circle this path creates (kilobots,
// lib/body.py Circle), so circle->m_p == (0,0) and b2Mul(xf, m_p) == xf.p exactly.
__device__ __forceinline__ void collide_circles(Manifold& m, float rA, Xf xfA, float rB, Xf xfB) {
  m.pointCount = 0;
  V2 pA = xmul(xfA, mk(0.0f, 0.0f));
  V2 pB = xmul(xfB, mk(0.0f, 0.0f));
  V2 d = pB - pA;
  float distSqr = dot(d, d);
  float radius = rA + rB;
  if (distSqr > radius * radius) return;
  m.type = MANIFOLD_CIRCLES;
  m.lpx = 0.0f; m.lpy = 0.0f;
  m.lnx = 0.0f; m.lny = 0.0f;
  m.pointCount = 1;
  m.px[0] = 0.0f; m.py[0] = 0.0f;
  m.id[0] = 0u;
}

__device__ __forceinline__ void collide_polygon_circle(Manifold& m, const ProxyConst* __restrict__ polyA, Xf xfA,
                                                    float circleRadius, Xf xfB) {
  m.pointCount = 0;
  V2 c = xmul(xfB, mk(0.0f, 0.0f));
  V2 cLocal = xmulT(xfA, c);
  int normalIndex = 0;
  float separation = -KB_MAXFLOAT;
  float radius = __ldg(&polyA->radius) + circleRadius;
  int vertexCount = __ldg(&polyA->count);
  for (int i = 0; i < vertexCount; ++i) {
    float s = dot(pnorm(polyA, i), cLocal - pvert(polyA, i));
    if (s > radius) return;
    if (s > separation) {
      separation = s;
      normalIndex = i;
    }
  }
  int vertIndex1 = normalIndex;
  int vertIndex2 = vertIndex1 + 1 < vertexCount ? vertIndex1 + 1 : 0;
  V2 v1 = pvert(polyA, vertIndex1);
  V2 v2 = pvert(polyA, vertIndex2);
  m.type = MANIFOLD_FACE_A;
  m.px[0] = 0.0f; m.py[0] = 0.0f;
  m.id[0] = 0u;
  if (separation < KB_EPS) {
    m.pointCount = 1;
    V2 n = pnorm(polyA, normalIndex);
    m.lnx = n.x; m.lny = n.y;
    V2 lp = 0.5f * (v1 + v2);
    m.lpx = lp.x; m.lpy = lp.y;
    return;
  }
  float u1 = dot(cLocal - v1, v2 - v1);
  float u2 = dot(cLocal - v2, v1 - v2);
  if (u1 <= 0.0f) {
    if (distsq(cLocal, v1) > radius * radius) return;
    m.pointCount = 1;
    V2 n = cLocal - v1;
    normalize(n);
    m.lnx = n.x; m.lny = n.y;
    m.lpx = v1.x; m.lpy = v1.y;
  } else if (u2 <= 0.0f) {
    if (distsq(cLocal, v2) > radius * radius) return;
    m.pointCount = 1;
    V2 n = cLocal - v2;
    normalize(n);
    m.lnx = n.x; m.lny = n.y;
    m.lpx = v2.x; m.lpy = v2.y;
  } else {
    V2 faceCenter = 0.5f * (v1 + v2);
    V2 n = pnorm(polyA, vertIndex1);
    float sep = dot(cLocal - faceCenter, n);
    if (sep > radius) return;
    m.pointCount = 1;
    m.lnx = n.x; m.lny = n.y;
    m.lpx = faceCenter.x; m.lpy = faceCenter.y;
  }
}

struct ClipV {
  V2 v;
  uint32_t id;
};

__device__ __forceinline__ int clip_segment(ClipV vOut[2], const ClipV vIn[2], V2 normal, float offset, int vertexIndexA) {
  int numOut = 0;
  float distance0 = dot(normal, vIn[0].v) - offset;
  float distance1 = dot(normal, vIn[1].v) - offset;
  if (distance0 <= 0.0f) vOut[numOut++] = vIn[0];
  if (distance1 <= 0.0f) vOut[numOut++] = vIn[1];
  if (distance0 * distance1 < 0.0f) {
    float interp = distance0 / (distance0 - distance1);
    vOut[numOut].v = vIn[0].v + interp * (vIn[1].v - vIn[0].v);
    // indexA = vertexIndexA, indexB = vIn[0].indexB, typeA = vertex(0), typeB = face(1)
    vOut[numOut].id = make_id((uint32_t)vertexIndexA, (vIn[0].id >> 8) & 0xFFu, 0u, 1u);
    ++numOut;
  }
  return numOut;
}

__device__ __forceinline__ float find_max_separation(int* edgeIndex, const ProxyConst* __restrict__ poly1, Xf xf1,
                                                     const ProxyConst* __restrict__ poly2, Xf xf2) {
  int count1 = __ldg(&poly1->count);
  int count2 = __ldg(&poly2->count);
  Xf xf = xmulT(xf2, xf1);
  int bestIndex = 0;
  float maxSeparation = -KB_MAXFLOAT;
  for (int i = 0; i < count1; ++i) {
    V2 n = rmul(xf.q, pnorm(poly1, i));
    V2 v1 = xmul(xf, pvert(poly1, i));
    float si = KB_MAXFLOAT;
    for (int j = 0; j < count2; ++j) {
      float sij = dot(n, pvert(poly2, j) - v1);
      if (sij < si) si = sij;
    }
    if (si > maxSeparation) {
      maxSeparation = si;
      bestIndex = i;
    }
  }
  *edgeIndex = bestIndex;
  return maxSeparation;
}

__device__ __noinline__ void collide_polygons(Manifold& m, const ProxyConst* __restrict__ polyA, Xf xfA,
                                              const ProxyConst* __restrict__ polyB, Xf xfB) {
  m.pointCount = 0;
  float totalRadius = __ldg(&polyA->radius) + __ldg(&polyB->radius);
  int edgeA = 0;
  float separationA = find_max_separation(&edgeA, polyA, xfA, polyB, xfB);
  if (separationA > totalRadius) return;
  int edgeB = 0;
  float separationB = find_max_separation(&edgeB, polyB, xfB, polyA, xfA);
  if (separationB > totalRadius) return;
  const ProxyConst* poly1;
  const ProxyConst* poly2;
  Xf xf1, xf2;
  int edge1;
  bool flip;
  const float k_tol = 0.1f * KB_LINEAR_SLOP;
  if (separationB > separationA + k_tol) {
    poly1 = polyB; poly2 = polyA; xf1 = xfB; xf2 = xfA; edge1 = edgeB;
    m.type = MANIFOLD_FACE_B;
    flip = true;
  } else {
    poly1 = polyA; poly2 = polyB; xf1 = xfA; xf2 = xfB; edge1 = edgeA;
    m.type = MANIFOLD_FACE_A;
    flip = false;
  }
  // b2FindIncidentEdge
  ClipV incidentEdge[2];
  {
    int count2 = __ldg(&poly2->count);
    V2 normal1 = rmulT(xf2.q, rmul(xf1.q, pnorm(poly1, edge1)));
    int index = 0;
    float minDot = KB_MAXFLOAT;
    for (int i = 0; i < count2; ++i) {
      float d = dot(normal1, pnorm(poly2, i));
      if (d < minDot) {
        minDot = d;
        index = i;
      }
    }
    int i1 = index;
    int i2 = i1 + 1 < count2 ? i1 + 1 : 0;
    incidentEdge[0].v = xmul(xf2, pvert(poly2, i1));
    incidentEdge[0].id = make_id((uint32_t)edge1, (uint32_t)i1, 1u, 0u);
    incidentEdge[1].v = xmul(xf2, pvert(poly2, i2));
    incidentEdge[1].id = make_id((uint32_t)edge1, (uint32_t)i2, 1u, 0u);
  }
  int count1 = __ldg(&poly1->count);
  int iv1 = edge1;
  int iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
  V2 v11 = pvert(poly1, iv1);
  V2 v12 = pvert(poly1, iv2);
  V2 localTangent = v12 - v11;
  normalize(localTangent);
  V2 localNormal = cross(localTangent, 1.0f);
  V2 planePoint = 0.5f * (v11 + v12);
  V2 tangent = rmul(xf1.q, localTangent);
  V2 normal = cross(tangent, 1.0f);
  v11 = xmul(xf1, v11);
  v12 = xmul(xf1, v12);
  float frontOffset = dot(normal, v11);
  float sideOffset1 = -dot(tangent, v11) + totalRadius;
  float sideOffset2 = dot(tangent, v12) + totalRadius;
  ClipV clipPoints1[2];
  ClipV clipPoints2[2];
  int np = clip_segment(clipPoints1, incidentEdge, -tangent, sideOffset1, iv1);
  if (np < 2) return;
  np = clip_segment(clipPoints2, clipPoints1, tangent, sideOffset2, iv2);
  if (np < 2) return;
  m.lnx = localNormal.x; m.lny = localNormal.y;
  m.lpx = planePoint.x; m.lpy = planePoint.y;
  int pointCount = 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float separation = dot(normal, clipPoints2[i].v) - frontOffset;
    if (separation <= totalRadius) {
      V2 lp = xmulT(xf2, clipPoints2[i].v);
      uint32_t id = clipPoints2[i].id;
      if (flip) id = swap_id(id);
      if (pointCount == 0) { m.px[0] = lp.x; m.py[0] = lp.y; m.id[0] = id; }
      else { m.px[1] = lp.x; m.py[1] = lp.y; m.id[1] = id; }
      ++pointCount;
    }
  }
  m.pointCount = pointCount;
}

// edge vertices of a chain child: vx/vy[0..3] = v0, v1, v2, v3
__device__ __forceinline__ void collide_edge_circle(Manifold& m, const ProxyConst* __restrict__ edgeA, Xf xfA,
                                                 float circleRadius, Xf xfB) {
  m.pointCount = 0;
  V2 Q = xmulT(xfA, xmul(xfB, mk(0.0f, 0.0f)));
  V2 A = pvert(edgeA, 1), B = pvert(edgeA, 2);
  V2 e = B - A;
  float u = dot(e, B - Q);
  float v = dot(e, Q - A);
  float radius = __ldg(&edgeA->radius) + circleRadius;
  m.px[0] = 0.0f; m.py[0] = 0.0f;
  if (v <= 0.0f) {
    V2 P = A;
    V2 d = Q - P;
    float dd = dot(d, d);
    if (dd > radius * radius) return;
    if (__ldg(&edgeA->has0)) {
      V2 A1 = pvert(edgeA, 0);
      V2 B1 = A;
      V2 e1 = B1 - A1;
      float u1 = dot(e1, B1 - Q);
      if (u1 > 0.0f) return;
    }
    m.pointCount = 1;
    m.type = MANIFOLD_CIRCLES;
    m.lnx = 0.0f; m.lny = 0.0f;
    m.lpx = P.x; m.lpy = P.y;
    m.id[0] = make_id(0u, 0u, 0u, 0u);
    return;
  }
  if (u <= 0.0f) {
    V2 P = B;
    V2 d = Q - P;
    float dd = dot(d, d);
    if (dd > radius * radius) return;
    if (__ldg(&edgeA->has3)) {
      V2 B2 = pvert(edgeA, 3);
      V2 A2 = B;
      V2 e2 = B2 - A2;
      float v2 = dot(e2, Q - A2);
      if (v2 > 0.0f) return;
    }
    m.pointCount = 1;
    m.type = MANIFOLD_CIRCLES;
    m.lnx = 0.0f; m.lny = 0.0f;
    m.lpx = P.x; m.lpy = P.y;
    m.id[0] = make_id(1u, 0u, 0u, 0u);
    return;
  }
  float den = dot(e, e);
  V2 P = (1.0f / den) * (u * A + v * B);
  V2 d = Q - P;
  float dd = dot(d, d);
  if (dd > radius * radius) return;
  V2 n = mk(-e.y, e.x);
  if (dot(n, Q - A) < 0.0f) n = mk(-n.x, -n.y);
  normalize(n);
  m.pointCount = 1;
  m.type = MANIFOLD_FACE_A;
  m.lnx = n.x; m.lny = n.y;
  m.lpx = A.x; m.lpy = A.y;
  m.id[0] = make_id(0u, 0u, 1u, 0u);
}

// b2EPCollider::Collide
__device__ __noinline__ void collide_edge_polygon(Manifold& m, const ProxyConst* __restrict__ edgeA, Xf xfA,
                                                  const ProxyConst* __restrict__ polyB, Xf xfB) {
  const Xf xf = xmulT(xfA, xfB);
  const V2 centroidB = xmul(xf, mk(__ldg(&polyB->cx), __ldg(&polyB->cy)));
  const V2 v0 = pvert(edgeA, 0), v1 = pvert(edgeA, 1), v2 = pvert(edgeA, 2), v3 = pvert(edgeA, 3);
  const bool hasVertex0 = __ldg(&edgeA->has0) != 0;
  const bool hasVertex3 = __ldg(&edgeA->has3) != 0;
  V2 edge1 = v2 - v1;
  normalize(edge1);
  const V2 normal1 = mk(edge1.y, -edge1.x);
  float offset1 = dot(normal1, centroidB - v1);
  float offset0 = 0.0f, offset2 = 0.0f;
  bool convex1 = false, convex2 = false;
  V2 normal0 = mk(0.0f, 0.0f), normal2 = mk(0.0f, 0.0f);
  if (hasVertex0) {
    V2 edge0 = v1 - v0;
    normalize(edge0);
    normal0 = mk(edge0.y, -edge0.x);
    convex1 = cross(edge0, edge1) >= 0.0f;
    offset0 = dot(normal0, centroidB - v0);
  }
  if (hasVertex3) {
    V2 edge2 = v3 - v2;
    normalize(edge2);
    normal2 = mk(edge2.y, -edge2.x);
    convex2 = cross(edge1, edge2) > 0.0f;
    offset2 = dot(normal2, centroidB - v2);
  }
  bool front;
  V2 normal, lowerLimit, upperLimit;
  if (hasVertex0 && hasVertex3) {
    if (convex1 && convex2) {
      front = offset0 >= 0.0f || offset1 >= 0.0f || offset2 >= 0.0f;
      if (front) { normal = normal1; lowerLimit = normal0; upperLimit = normal2; }
      else { normal = -normal1; lowerLimit = -normal1; upperLimit = -normal1; }
    } else if (convex1) {
      front = offset0 >= 0.0f || (offset1 >= 0.0f && offset2 >= 0.0f);
      if (front) { normal = normal1; lowerLimit = normal0; upperLimit = normal1; }
      else { normal = -normal1; lowerLimit = -normal2; upperLimit = -normal1; }
    } else if (convex2) {
      front = offset2 >= 0.0f || (offset0 >= 0.0f && offset1 >= 0.0f);
      if (front) { normal = normal1; lowerLimit = normal1; upperLimit = normal2; }
      else { normal = -normal1; lowerLimit = -normal1; upperLimit = -normal0; }
    } else {
      front = offset0 >= 0.0f && offset1 >= 0.0f && offset2 >= 0.0f;
      if (front) { normal = normal1; lowerLimit = normal1; upperLimit = normal1; }
      else { normal = -normal1; lowerLimit = -normal2; upperLimit = -normal0; }
    }
  } else if (hasVertex0) {
    if (convex1) {
      front = offset0 >= 0.0f || offset1 >= 0.0f;
      if (front) { normal = normal1; lowerLimit = normal0; upperLimit = -normal1; }
      else { normal = -normal1; lowerLimit = normal1; upperLimit = -normal1; }
    } else {
      front = offset0 >= 0.0f && offset1 >= 0.0f;
      if (front) { normal = normal1; lowerLimit = normal1; upperLimit = -normal1; }
      else { normal = -normal1; lowerLimit = normal1; upperLimit = -normal0; }
    }
  } else if (hasVertex3) {
    if (convex2) {
      front = offset1 >= 0.0f || offset2 >= 0.0f;
      if (front) { normal = normal1; lowerLimit = -normal1; upperLimit = normal2; }
      else { normal = -normal1; lowerLimit = -normal1; upperLimit = normal1; }
    } else {
      front = offset1 >= 0.0f && offset2 >= 0.0f;
      if (front) { normal = normal1; lowerLimit = -normal1; upperLimit = normal1; }
      else { normal = -normal1; lowerLimit = -normal2; upperLimit = normal1; }
    }
  } else {
    front = offset1 >= 0.0f;
    if (front) { normal = normal1; lowerLimit = -normal1; upperLimit = -normal1; }
    else { normal = -normal1; lowerLimit = normal1; upperLimit = normal1; }
  }
  const int countB = __ldg(&polyB->count);
  const float radius = 2.0f * KB_POLYGON_RADIUS;
  m.pointCount = 0;
  // ComputeEdgeSeparation
  float edgeSeparation = KB_MAXFLOAT;
  for (int i = 0; i < countB; ++i) {
    float s = dot(normal, xmul(xf, pvert(polyB, i)) - v1);
    if (s < edgeSeparation) edgeSeparation = s;
  }
  if (edgeSeparation > radius) return;
  // ComputePolygonSeparation
  int polyType = 0;  // 0 unknown, 2 edgeB
  int polyIndex = -1;
  float polySeparation = -KB_MAXFLOAT;
  {
    V2 perp = mk(-normal.y, normal.x);
    for (int i = 0; i < countB; ++i) {
      V2 n = -rmul(xf.q, pnorm(polyB, i));
      V2 vb = xmul(xf, pvert(polyB, i));
      float s1 = dot(n, vb - v1);
      float s2 = dot(n, vb - v2);
      float s = b2min(s1, s2);
      if (s > radius) {
        polyType = 2;
        polyIndex = i;
        polySeparation = s;
        break;
      }
      if (dot(n, perp) >= 0.0f) {
        if (dot(n - upperLimit, normal) < -KB_ANGULAR_SLOP) continue;
      } else {
        if (dot(n - lowerLimit, normal) < -KB_ANGULAR_SLOP) continue;
      }
      if (s > polySeparation) {
        polyType = 2;
        polyIndex = i;
        polySeparation = s;
      }
    }
  }
  if (polyType != 0 && polySeparation > radius) return;
  const float k_relativeTol = 0.98f;
  const float k_absoluteTol = 0.001f;
  bool usePolygonAxis;
  if (polyType == 0) usePolygonAxis = false;
  else if (polySeparation > k_relativeTol * edgeSeparation + k_absoluteTol) usePolygonAxis = true;
  else usePolygonAxis = false;

  ClipV ie[2];
  int rf_i1, rf_i2;
  V2 rf_v1, rf_v2, rf_normal;
  if (!usePolygonAxis) {
    m.type = MANIFOLD_FACE_A;
    int bestIndex = 0;
    float bestValue = dot(normal, rmul(xf.q, pnorm(polyB, 0)));
    for (int i = 1; i < countB; ++i) {
      float value = dot(normal, rmul(xf.q, pnorm(polyB, i)));
      if (value < bestValue) {
        bestValue = value;
        bestIndex = i;
      }
    }
    int i1 = bestIndex;
    int i2 = i1 + 1 < countB ? i1 + 1 : 0;
    ie[0].v = xmul(xf, pvert(polyB, i1));
    ie[0].id = make_id(0u, (uint32_t)i1, 1u, 0u);
    ie[1].v = xmul(xf, pvert(polyB, i2));
    ie[1].id = make_id(0u, (uint32_t)i2, 1u, 0u);
    if (front) {
      rf_i1 = 0; rf_i2 = 1; rf_v1 = v1; rf_v2 = v2; rf_normal = normal1;
    } else {
      rf_i1 = 1; rf_i2 = 0; rf_v1 = v2; rf_v2 = v1; rf_normal = -normal1;
    }
  } else {
    m.type = MANIFOLD_FACE_B;
    ie[0].v = v1;
    ie[0].id = make_id(0u, (uint32_t)polyIndex, 0u, 1u);
    ie[1].v = v2;
    ie[1].id = make_id(0u, (uint32_t)polyIndex, 0u, 1u);
    rf_i1 = polyIndex;
    rf_i2 = rf_i1 + 1 < countB ? rf_i1 + 1 : 0;
    rf_v1 = xmul(xf, pvert(polyB, rf_i1));
    rf_v2 = xmul(xf, pvert(polyB, rf_i2));
    rf_normal = rmul(xf.q, pnorm(polyB, rf_i1));
  }
  V2 sideNormal1 = mk(rf_normal.y, -rf_normal.x);
  V2 sideNormal2 = -sideNormal1;
  float sideOffset1 = dot(sideNormal1, rf_v1);
  float sideOffset2 = dot(sideNormal2, rf_v2);
  ClipV clipPoints1[2];
  ClipV clipPoints2[2];
  int np = clip_segment(clipPoints1, ie, sideNormal1, sideOffset1, rf_i1);
  if (np < 2) return;
  np = clip_segment(clipPoints2, clipPoints1, sideNormal2, sideOffset2, rf_i2);
  if (np < 2) return;
  if (!usePolygonAxis) {
    m.lnx = rf_normal.x; m.lny = rf_normal.y;
    m.lpx = rf_v1.x; m.lpy = rf_v1.y;
  } else {
    V2 ln = pnorm(polyB, rf_i1);
    V2 lp = pvert(polyB, rf_i1);
    m.lnx = ln.x; m.lny = ln.y;
    m.lpx = lp.x; m.lpy = lp.y;
  }
  int pointCount = 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float separation = dot(rf_normal, clipPoints2[i].v - rf_v1);
    if (separation <= radius) {
      V2 lp;
      uint32_t id;
      if (!usePolygonAxis) {
        lp = xmulT(xf, clipPoints2[i].v);
        id = clipPoints2[i].id;
      } else {
        lp = clipPoints2[i].v;
        id = swap_id(clipPoints2[i].id);
      }
      if (pointCount == 0) { m.px[0] = lp.x; m.py[0] = lp.y; m.id[0] = id; }
      else { m.px[1] = lp.x; m.py[1] = lp.y; m.id[1] = id; }
      ++pointCount;
    }
  }
  m.pointCount = pointCount;
}

}  // namespace kb
