// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see kbo_math.h).
// Narrowphase: restates b2CollideCircle.cpp, b2CollidePolygon.cpp (2.3.1 brute-force
// b2FindMaxSeparation), b2CollideEdge.cpp (b2CollideEdgeAndCircle, b2EPCollider) and
// b2ClipSegmentToLine of Box2D 2.3.x -- SURVEY.md Appendix B.5.1/B.5.2.
#include "kbo_world.h"

namespace kbo {

void CollideCircles(Manifold* manifold, const Shape* circleA, const Xf& xfA, const Shape* circleB,
                    const Xf& xfB) {
  manifold->pointCount = 0;
  Vec2 pA = Mul(xfA, circleA->p);
  Vec2 pB = Mul(xfB, circleB->p);
  Vec2 d = pB - pA;
  float distSqr = Dot(d, d);
  float rA = circleA->radius, rB = circleB->radius;
  float radius = rA + rB;
  if (distSqr > radius * radius) return;
  manifold->type = kManifoldCircles;
  manifold->localPoint = circleA->p;
  manifold->localNormal.SetZero();
  manifold->pointCount = 1;
  manifold->points[0].localPoint = circleB->p;
  manifold->points[0].id.key = 0;
}

void CollidePolygonAndCircle(Manifold* manifold, const Shape* polygonA, const Xf& xfA,
                             const Shape* circleB, const Xf& xfB) {
  manifold->pointCount = 0;
  Vec2 c = Mul(xfB, circleB->p);
  Vec2 cLocal = MulT(xfA, c);
  int normalIndex = 0;
  float separation = -kMaxFloat;
  float radius = polygonA->radius + circleB->radius;
  int vertexCount = polygonA->count;
  const Vec2* vertices = polygonA->vertices;
  const Vec2* normals = polygonA->normals;
  for (int i = 0; i < vertexCount; ++i) {
    float s = Dot(normals[i], cLocal - vertices[i]);
    if (s > radius) return;
    if (s > separation) {
      separation = s;
      normalIndex = i;
    }
  }
  int vertIndex1 = normalIndex;
  int vertIndex2 = vertIndex1 + 1 < vertexCount ? vertIndex1 + 1 : 0;
  Vec2 v1 = vertices[vertIndex1];
  Vec2 v2 = vertices[vertIndex2];
  if (separation < kEpsilon) {
    manifold->pointCount = 1;
    manifold->type = kManifoldFaceA;
    manifold->localNormal = normals[normalIndex];
    manifold->localPoint = 0.5f * (v1 + v2);
    manifold->points[0].localPoint = circleB->p;
    manifold->points[0].id.key = 0;
    return;
  }
  float u1 = Dot(cLocal - v1, v2 - v1);
  float u2 = Dot(cLocal - v2, v1 - v2);
  if (u1 <= 0.0f) {
    if (DistanceSquared(cLocal, v1) > radius * radius) return;
    manifold->pointCount = 1;
    manifold->type = kManifoldFaceA;
    manifold->localNormal = cLocal - v1;
    manifold->localNormal.Normalize();
    manifold->localPoint = v1;
    manifold->points[0].localPoint = circleB->p;
    manifold->points[0].id.key = 0;
  } else if (u2 <= 0.0f) {
    if (DistanceSquared(cLocal, v2) > radius * radius) return;
    manifold->pointCount = 1;
    manifold->type = kManifoldFaceA;
    manifold->localNormal = cLocal - v2;
    manifold->localNormal.Normalize();
    manifold->localPoint = v2;
    manifold->points[0].localPoint = circleB->p;
    manifold->points[0].id.key = 0;
  } else {
    Vec2 faceCenter = 0.5f * (v1 + v2);
    float sep = Dot(cLocal - faceCenter, normals[vertIndex1]);
    if (sep > radius) return;
    manifold->pointCount = 1;
    manifold->type = kManifoldFaceA;
    manifold->localNormal = normals[vertIndex1];
    manifold->localPoint = faceCenter;
    manifold->points[0].localPoint = circleB->p;
    manifold->points[0].id.key = 0;
  }
}

static int ClipSegmentToLine(ClipVertex vOut[2], const ClipVertex vIn[2], const Vec2& normal,
                             float offset, int vertexIndexA) {
  int numOut = 0;
  float distance0 = Dot(normal, vIn[0].v) - offset;
  float distance1 = Dot(normal, vIn[1].v) - offset;
  if (distance0 <= 0.0f) vOut[numOut++] = vIn[0];
  if (distance1 <= 0.0f) vOut[numOut++] = vIn[1];
  if (distance0 * distance1 < 0.0f) {
    float interp = distance0 / (distance0 - distance1);
    vOut[numOut].v = vIn[0].v + interp * (vIn[1].v - vIn[0].v);
    vOut[numOut].id.cf.indexA = (uint8_t)vertexIndexA;
    vOut[numOut].id.cf.indexB = vIn[0].id.cf.indexB;
    vOut[numOut].id.cf.typeA = kFeatureVertex;
    vOut[numOut].id.cf.typeB = kFeatureFace;
    ++numOut;
  }
  return numOut;
}

static float FindMaxSeparation(int* edgeIndex, const Shape* poly1, const Xf& xf1, const Shape* poly2,
                               const Xf& xf2) {
  int count1 = poly1->count;
  int count2 = poly2->count;
  const Vec2* n1s = poly1->normals;
  const Vec2* v1s = poly1->vertices;
  const Vec2* v2s = poly2->vertices;
  Xf xf = MulT(xf2, xf1);
  int bestIndex = 0;
  float maxSeparation = -kMaxFloat;
  for (int i = 0; i < count1; ++i) {
    Vec2 n = Mul(xf.q, n1s[i]);
    Vec2 v1 = Mul(xf, v1s[i]);
    float si = kMaxFloat;
    for (int j = 0; j < count2; ++j) {
      float sij = Dot(n, v2s[j] - v1);
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

static void FindIncidentEdge(ClipVertex c[2], const Shape* poly1, const Xf& xf1, int edge1,
                             const Shape* poly2, const Xf& xf2) {
  const Vec2* normals1 = poly1->normals;
  int count2 = poly2->count;
  const Vec2* vertices2 = poly2->vertices;
  const Vec2* normals2 = poly2->normals;
  Vec2 normal1 = MulT(xf2.q, Mul(xf1.q, normals1[edge1]));
  int index = 0;
  float minDot = kMaxFloat;
  for (int i = 0; i < count2; ++i) {
    float dot = Dot(normal1, normals2[i]);
    if (dot < minDot) {
      minDot = dot;
      index = i;
    }
  }
  int i1 = index;
  int i2 = i1 + 1 < count2 ? i1 + 1 : 0;
  c[0].v = Mul(xf2, vertices2[i1]);
  c[0].id.cf.indexA = (uint8_t)edge1;
  c[0].id.cf.indexB = (uint8_t)i1;
  c[0].id.cf.typeA = kFeatureFace;
  c[0].id.cf.typeB = kFeatureVertex;
  c[1].v = Mul(xf2, vertices2[i2]);
  c[1].id.cf.indexA = (uint8_t)edge1;
  c[1].id.cf.indexB = (uint8_t)i2;
  c[1].id.cf.typeA = kFeatureFace;
  c[1].id.cf.typeB = kFeatureVertex;
}

void CollidePolygons(Manifold* manifold, const Shape* polyA, const Xf& xfA, const Shape* polyB,
                     const Xf& xfB) {
  manifold->pointCount = 0;
  float totalRadius = polyA->radius + polyB->radius;
  int edgeA = 0;
  float separationA = FindMaxSeparation(&edgeA, polyA, xfA, polyB, xfB);
  if (separationA > totalRadius) return;
  int edgeB = 0;
  float separationB = FindMaxSeparation(&edgeB, polyB, xfB, polyA, xfA);
  if (separationB > totalRadius) return;
  const Shape* poly1;
  const Shape* poly2;
  Xf xf1, xf2;
  int edge1;
  uint8_t flip;
  const float k_tol = 0.1f * kLinearSlop;
  if (separationB > separationA + k_tol) {
    poly1 = polyB;
    poly2 = polyA;
    xf1 = xfB;
    xf2 = xfA;
    edge1 = edgeB;
    manifold->type = kManifoldFaceB;
    flip = 1;
  } else {
    poly1 = polyA;
    poly2 = polyB;
    xf1 = xfA;
    xf2 = xfB;
    edge1 = edgeA;
    manifold->type = kManifoldFaceA;
    flip = 0;
  }
  ClipVertex incidentEdge[2];
  FindIncidentEdge(incidentEdge, poly1, xf1, edge1, poly2, xf2);
  int count1 = poly1->count;
  const Vec2* vertices1 = poly1->vertices;
  int iv1 = edge1;
  int iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
  Vec2 v11 = vertices1[iv1];
  Vec2 v12 = vertices1[iv2];
  Vec2 localTangent = v12 - v11;
  localTangent.Normalize();
  Vec2 localNormal = Cross(localTangent, 1.0f);
  Vec2 planePoint = 0.5f * (v11 + v12);
  Vec2 tangent = Mul(xf1.q, localTangent);
  Vec2 normal = Cross(tangent, 1.0f);
  v11 = Mul(xf1, v11);
  v12 = Mul(xf1, v12);
  float frontOffset = Dot(normal, v11);
  float sideOffset1 = -Dot(tangent, v11) + totalRadius;
  float sideOffset2 = Dot(tangent, v12) + totalRadius;
  ClipVertex clipPoints1[2];
  ClipVertex clipPoints2[2];
  int np;
  np = ClipSegmentToLine(clipPoints1, incidentEdge, -tangent, sideOffset1, iv1);
  if (np < 2) return;
  np = ClipSegmentToLine(clipPoints2, clipPoints1, tangent, sideOffset2, iv2);
  if (np < 2) return;
  manifold->localNormal = localNormal;
  manifold->localPoint = planePoint;
  int pointCount = 0;
  for (int i = 0; i < kMaxManifoldPoints; ++i) {
    float separation = Dot(normal, clipPoints2[i].v) - frontOffset;
    if (separation <= totalRadius) {
      ManifoldPoint* cp = manifold->points + pointCount;
      cp->localPoint = MulT(xf2, clipPoints2[i].v);
      cp->id = clipPoints2[i].id;
      if (flip) {
        ContactFeature cf = cp->id.cf;
        cp->id.cf.indexA = cf.indexB;
        cp->id.cf.indexB = cf.indexA;
        cp->id.cf.typeA = cf.typeB;
        cp->id.cf.typeB = cf.typeA;
      }
      ++pointCount;
    }
  }
  manifold->pointCount = pointCount;
}

void CollideEdgeAndCircle(Manifold* manifold, const Shape* edgeA, const Xf& xfA, const Shape* circleB,
                          const Xf& xfB) {
  manifold->pointCount = 0;
  Vec2 Q = MulT(xfA, Mul(xfB, circleB->p));
  Vec2 A = edgeA->v1, B = edgeA->v2;
  Vec2 e = B - A;
  float u = Dot(e, B - Q);
  float v = Dot(e, Q - A);
  float radius = edgeA->radius + circleB->radius;
  ContactFeature cf;
  cf.indexB = 0;
  cf.typeB = kFeatureVertex;
  if (v <= 0.0f) {
    Vec2 P = A;
    Vec2 d = Q - P;
    float dd = Dot(d, d);
    if (dd > radius * radius) return;
    if (edgeA->hasVertex0) {
      Vec2 A1 = edgeA->v0;
      Vec2 B1 = A;
      Vec2 e1 = B1 - A1;
      float u1 = Dot(e1, B1 - Q);
      if (u1 > 0.0f) return;
    }
    cf.indexA = 0;
    cf.typeA = kFeatureVertex;
    manifold->pointCount = 1;
    manifold->type = kManifoldCircles;
    manifold->localNormal.SetZero();
    manifold->localPoint = P;
    manifold->points[0].id.key = 0;
    manifold->points[0].id.cf = cf;
    manifold->points[0].localPoint = circleB->p;
    return;
  }
  if (u <= 0.0f) {
    Vec2 P = B;
    Vec2 d = Q - P;
    float dd = Dot(d, d);
    if (dd > radius * radius) return;
    if (edgeA->hasVertex3) {
      Vec2 B2 = edgeA->v3;
      Vec2 A2 = B;
      Vec2 e2 = B2 - A2;
      float v2 = Dot(e2, Q - A2);
      if (v2 > 0.0f) return;
    }
    cf.indexA = 1;
    cf.typeA = kFeatureVertex;
    manifold->pointCount = 1;
    manifold->type = kManifoldCircles;
    manifold->localNormal.SetZero();
    manifold->localPoint = P;
    manifold->points[0].id.key = 0;
    manifold->points[0].id.cf = cf;
    manifold->points[0].localPoint = circleB->p;
    return;
  }
  float den = Dot(e, e);
  Vec2 P = (1.0f / den) * (u * A + v * B);
  Vec2 d = Q - P;
  float dd = Dot(d, d);
  if (dd > radius * radius) return;
  Vec2 n(-e.y, e.x);
  if (Dot(n, Q - A) < 0.0f) n.Set(-n.x, -n.y);
  n.Normalize();
  cf.indexA = 0;
  cf.typeA = kFeatureFace;
  manifold->pointCount = 1;
  manifold->type = kManifoldFaceA;
  manifold->localNormal = n;
  manifold->localPoint = A;
  manifold->points[0].id.key = 0;
  manifold->points[0].id.cf = cf;
  manifold->points[0].localPoint = circleB->p;
}

// --- b2EPCollider ------------------------------------------------------------------------------
namespace {
struct EPAxis {
  enum Type { kUnknown, kEdgeA, kEdgeB };
  Type type;
  int index;
  float separation;
};
struct TempPolygon {
  Vec2 vertices[kMaxPolygonVertices];
  Vec2 normals[kMaxPolygonVertices];
  int count;
};
struct ReferenceFace {
  int i1, i2;
  Vec2 v1, v2;
  Vec2 normal;
  Vec2 sideNormal1;
  float sideOffset1;
  Vec2 sideNormal2;
  float sideOffset2;
};
struct EPCollider {
  TempPolygon m_polygonB;
  Xf m_xf;
  Vec2 m_centroidB;
  Vec2 m_v0, m_v1, m_v2, m_v3;
  Vec2 m_normal0, m_normal1, m_normal2;
  Vec2 m_normal;
  Vec2 m_lowerLimit, m_upperLimit;
  float m_radius;
  bool m_front;

  EPAxis ComputeEdgeSeparation() {
    EPAxis axis;
    axis.type = EPAxis::kEdgeA;
    axis.index = m_front ? 0 : 1;
    axis.separation = FLT_MAX;
    for (int i = 0; i < m_polygonB.count; ++i) {
      float s = Dot(m_normal, m_polygonB.vertices[i] - m_v1);
      if (s < axis.separation) axis.separation = s;
    }
    return axis;
  }

  EPAxis ComputePolygonSeparation() {
    EPAxis axis;
    axis.type = EPAxis::kUnknown;
    axis.index = -1;
    axis.separation = -FLT_MAX;
    Vec2 perp(-m_normal.y, m_normal.x);
    for (int i = 0; i < m_polygonB.count; ++i) {
      Vec2 n = -m_polygonB.normals[i];
      float s1 = Dot(n, m_polygonB.vertices[i] - m_v1);
      float s2 = Dot(n, m_polygonB.vertices[i] - m_v2);
      float s = Min(s1, s2);
      if (s > m_radius) {
        axis.type = EPAxis::kEdgeB;
        axis.index = i;
        axis.separation = s;
        return axis;
      }
      if (Dot(n, perp) >= 0.0f) {
        if (Dot(n - m_upperLimit, m_normal) < -kAngularSlop) continue;
      } else {
        if (Dot(n - m_lowerLimit, m_normal) < -kAngularSlop) continue;
      }
      if (s > axis.separation) {
        axis.type = EPAxis::kEdgeB;
        axis.index = i;
        axis.separation = s;
      }
    }
    return axis;
  }

  void Collide(Manifold* manifold, const Shape* edgeA, const Xf& xfA, const Shape* polygonB,
               const Xf& xfB) {
    m_xf = MulT(xfA, xfB);
    m_centroidB = Mul(m_xf, polygonB->centroid);
    m_v0 = edgeA->v0;
    m_v1 = edgeA->v1;
    m_v2 = edgeA->v2;
    m_v3 = edgeA->v3;
    bool hasVertex0 = edgeA->hasVertex0;
    bool hasVertex3 = edgeA->hasVertex3;
    Vec2 edge1 = m_v2 - m_v1;
    edge1.Normalize();
    m_normal1.Set(edge1.y, -edge1.x);
    float offset1 = Dot(m_normal1, m_centroidB - m_v1);
    float offset0 = 0.0f, offset2 = 0.0f;
    bool convex1 = false, convex2 = false;
    if (hasVertex0) {
      Vec2 edge0 = m_v1 - m_v0;
      edge0.Normalize();
      m_normal0.Set(edge0.y, -edge0.x);
      convex1 = Cross(edge0, edge1) >= 0.0f;
      offset0 = Dot(m_normal0, m_centroidB - m_v0);
    }
    if (hasVertex3) {
      Vec2 edge2 = m_v3 - m_v2;
      edge2.Normalize();
      m_normal2.Set(edge2.y, -edge2.x);
      convex2 = Cross(edge1, edge2) > 0.0f;
      offset2 = Dot(m_normal2, m_centroidB - m_v2);
    }
    if (hasVertex0 && hasVertex3) {
      if (convex1 && convex2) {
        m_front = offset0 >= 0.0f || offset1 >= 0.0f || offset2 >= 0.0f;
        if (m_front) { m_normal = m_normal1; m_lowerLimit = m_normal0; m_upperLimit = m_normal2; }
        else { m_normal = -m_normal1; m_lowerLimit = -m_normal1; m_upperLimit = -m_normal1; }
      } else if (convex1) {
        m_front = offset0 >= 0.0f || (offset1 >= 0.0f && offset2 >= 0.0f);
        if (m_front) { m_normal = m_normal1; m_lowerLimit = m_normal0; m_upperLimit = m_normal1; }
        else { m_normal = -m_normal1; m_lowerLimit = -m_normal2; m_upperLimit = -m_normal1; }
      } else if (convex2) {
        m_front = offset2 >= 0.0f || (offset0 >= 0.0f && offset1 >= 0.0f);
        if (m_front) { m_normal = m_normal1; m_lowerLimit = m_normal1; m_upperLimit = m_normal2; }
        else { m_normal = -m_normal1; m_lowerLimit = -m_normal1; m_upperLimit = -m_normal0; }
      } else {
        m_front = offset0 >= 0.0f && offset1 >= 0.0f && offset2 >= 0.0f;
        if (m_front) { m_normal = m_normal1; m_lowerLimit = m_normal1; m_upperLimit = m_normal1; }
        else { m_normal = -m_normal1; m_lowerLimit = -m_normal2; m_upperLimit = -m_normal0; }
      }
    } else if (hasVertex0) {
      if (convex1) {
        m_front = offset0 >= 0.0f || offset1 >= 0.0f;
        if (m_front) { m_normal = m_normal1; m_lowerLimit = m_normal0; m_upperLimit = -m_normal1; }
        else { m_normal = -m_normal1; m_lowerLimit = m_normal1; m_upperLimit = -m_normal1; }
      } else {
        m_front = offset0 >= 0.0f && offset1 >= 0.0f;
        if (m_front) { m_normal = m_normal1; m_lowerLimit = m_normal1; m_upperLimit = -m_normal1; }
        else { m_normal = -m_normal1; m_lowerLimit = m_normal1; m_upperLimit = -m_normal0; }
      }
    } else if (hasVertex3) {
      if (convex2) {
        m_front = offset1 >= 0.0f || offset2 >= 0.0f;
        if (m_front) { m_normal = m_normal1; m_lowerLimit = -m_normal1; m_upperLimit = m_normal2; }
        else { m_normal = -m_normal1; m_lowerLimit = -m_normal1; m_upperLimit = m_normal1; }
      } else {
        m_front = offset1 >= 0.0f && offset2 >= 0.0f;
        if (m_front) { m_normal = m_normal1; m_lowerLimit = -m_normal1; m_upperLimit = m_normal1; }
        else { m_normal = -m_normal1; m_lowerLimit = -m_normal2; m_upperLimit = m_normal1; }
      }
    } else {
      m_front = offset1 >= 0.0f;
      if (m_front) { m_normal = m_normal1; m_lowerLimit = -m_normal1; m_upperLimit = -m_normal1; }
      else { m_normal = -m_normal1; m_lowerLimit = m_normal1; m_upperLimit = m_normal1; }
    }
    m_polygonB.count = polygonB->count;
    for (int i = 0; i < polygonB->count; ++i) {
      m_polygonB.vertices[i] = Mul(m_xf, polygonB->vertices[i]);
      m_polygonB.normals[i] = Mul(m_xf.q, polygonB->normals[i]);
    }
    m_radius = 2.0f * kPolygonRadius;
    manifold->pointCount = 0;
    EPAxis edgeAxis = ComputeEdgeSeparation();
    if (edgeAxis.type == EPAxis::kUnknown) return;
    if (edgeAxis.separation > m_radius) return;
    EPAxis polygonAxis = ComputePolygonSeparation();
    if (polygonAxis.type != EPAxis::kUnknown && polygonAxis.separation > m_radius) return;
    const float k_relativeTol = 0.98f;
    const float k_absoluteTol = 0.001f;
    EPAxis primaryAxis;
    if (polygonAxis.type == EPAxis::kUnknown) primaryAxis = edgeAxis;
    else if (polygonAxis.separation > k_relativeTol * edgeAxis.separation + k_absoluteTol) primaryAxis = polygonAxis;
    else primaryAxis = edgeAxis;
    ClipVertex ie[2];
    ReferenceFace rf;
    if (primaryAxis.type == EPAxis::kEdgeA) {
      manifold->type = kManifoldFaceA;
      int bestIndex = 0;
      float bestValue = Dot(m_normal, m_polygonB.normals[0]);
      for (int i = 1; i < m_polygonB.count; ++i) {
        float value = Dot(m_normal, m_polygonB.normals[i]);
        if (value < bestValue) {
          bestValue = value;
          bestIndex = i;
        }
      }
      int i1 = bestIndex;
      int i2 = i1 + 1 < m_polygonB.count ? i1 + 1 : 0;
      ie[0].v = m_polygonB.vertices[i1];
      ie[0].id.cf.indexA = 0;
      ie[0].id.cf.indexB = (uint8_t)i1;
      ie[0].id.cf.typeA = kFeatureFace;
      ie[0].id.cf.typeB = kFeatureVertex;
      ie[1].v = m_polygonB.vertices[i2];
      ie[1].id.cf.indexA = 0;
      ie[1].id.cf.indexB = (uint8_t)i2;
      ie[1].id.cf.typeA = kFeatureFace;
      ie[1].id.cf.typeB = kFeatureVertex;
      if (m_front) {
        rf.i1 = 0; rf.i2 = 1; rf.v1 = m_v1; rf.v2 = m_v2; rf.normal = m_normal1;
      } else {
        rf.i1 = 1; rf.i2 = 0; rf.v1 = m_v2; rf.v2 = m_v1; rf.normal = -m_normal1;
      }
    } else {
      manifold->type = kManifoldFaceB;
      ie[0].v = m_v1;
      ie[0].id.cf.indexA = 0;
      ie[0].id.cf.indexB = (uint8_t)primaryAxis.index;
      ie[0].id.cf.typeA = kFeatureVertex;
      ie[0].id.cf.typeB = kFeatureFace;
      ie[1].v = m_v2;
      ie[1].id.cf.indexA = 0;
      ie[1].id.cf.indexB = (uint8_t)primaryAxis.index;
      ie[1].id.cf.typeA = kFeatureVertex;
      ie[1].id.cf.typeB = kFeatureFace;
      rf.i1 = primaryAxis.index;
      rf.i2 = rf.i1 + 1 < m_polygonB.count ? rf.i1 + 1 : 0;
      rf.v1 = m_polygonB.vertices[rf.i1];
      rf.v2 = m_polygonB.vertices[rf.i2];
      rf.normal = m_polygonB.normals[rf.i1];
    }
    rf.sideNormal1.Set(rf.normal.y, -rf.normal.x);
    rf.sideNormal2 = -rf.sideNormal1;
    rf.sideOffset1 = Dot(rf.sideNormal1, rf.v1);
    rf.sideOffset2 = Dot(rf.sideNormal2, rf.v2);
    ClipVertex clipPoints1[2];
    ClipVertex clipPoints2[2];
    int np;
    np = ClipSegmentToLine(clipPoints1, ie, rf.sideNormal1, rf.sideOffset1, rf.i1);
    if (np < kMaxManifoldPoints) return;
    np = ClipSegmentToLine(clipPoints2, clipPoints1, rf.sideNormal2, rf.sideOffset2, rf.i2);
    if (np < kMaxManifoldPoints) return;
    if (primaryAxis.type == EPAxis::kEdgeA) {
      manifold->localNormal = rf.normal;
      manifold->localPoint = rf.v1;
    } else {
      manifold->localNormal = polygonB->normals[rf.i1];
      manifold->localPoint = polygonB->vertices[rf.i1];
    }
    int pointCount = 0;
    for (int i = 0; i < kMaxManifoldPoints; ++i) {
      float separation = Dot(rf.normal, clipPoints2[i].v - rf.v1);
      if (separation <= m_radius) {
        ManifoldPoint* cp = manifold->points + pointCount;
        if (primaryAxis.type == EPAxis::kEdgeA) {
          cp->localPoint = MulT(m_xf, clipPoints2[i].v);
          cp->id = clipPoints2[i].id;
        } else {
          cp->localPoint = clipPoints2[i].v;
          cp->id.cf.typeA = clipPoints2[i].id.cf.typeB;
          cp->id.cf.typeB = clipPoints2[i].id.cf.typeA;
          cp->id.cf.indexA = clipPoints2[i].id.cf.indexB;
          cp->id.cf.indexB = clipPoints2[i].id.cf.indexA;
        }
        ++pointCount;
      }
    }
    manifold->pointCount = pointCount;
  }
};
}  // namespace

void CollideEdgeAndPolygon(Manifold* manifold, const Shape* edgeA, const Xf& xfA, const Shape* polyB,
                           const Xf& xfB) {
  EPCollider collider;
  collider.Collide(manifold, edgeA, xfA, polyB, xfB);
}

}  // namespace kbo
