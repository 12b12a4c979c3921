// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see kbo_math.h).
// Restates b2ContactSolver.cpp of Box2D 2.3.x: constructor,
// InitializeVelocityConstraints (+ b2WorldManifold::Initialize), WarmStart,
// SolveVelocityConstraints (friction rows, 1-point normal row, 2-point block LCP), StoreImpulses,
// SolvePositionConstraints / SolveTOIPositionConstraints (+ b2PositionSolverManifold).
#include "kbo_world.h"

namespace kbo {

ContactSolver::ContactSolver(const std::vector<Contact*>& contacts_, std::vector<Position>* positions_,
                             std::vector<Velocity>* velocities_, float dtRatio, bool warmStarting)
    : contacts(contacts_), positions(*positions_), velocities(*velocities_) {
  const int count = (int)contacts.size();
  vcs.resize(count);
  pcs.resize(count);
  for (int i = 0; i < count; ++i) {
    Contact* contact = contacts[i];
    Fixture* fixtureA = contact->fixtureA;
    Fixture* fixtureB = contact->fixtureB;
    float radiusA = fixtureA->shape.radius;
    float radiusB = fixtureB->shape.radius;
    Body* bodyA = fixtureA->body;
    Body* bodyB = fixtureB->body;
    Manifold* manifold = &contact->manifold;
    int pointCount = manifold->pointCount;

    ContactVelocityConstraint* vc = &vcs[i];
    vc->friction = contact->friction;
    vc->restitution = contact->restitution;
    vc->indexA = bodyA->islandIndex;
    vc->indexB = bodyB->islandIndex;
    vc->invMassA = bodyA->invMass;
    vc->invMassB = bodyB->invMass;
    vc->invIA = bodyA->invI;
    vc->invIB = bodyB->invI;
    vc->contactIndex = i;
    vc->pointCount = pointCount;
    vc->K.SetZero();
    vc->normalMass.SetZero();

    ContactPositionConstraint* pc = &pcs[i];
    pc->indexA = bodyA->islandIndex;
    pc->indexB = bodyB->islandIndex;
    pc->invMassA = bodyA->invMass;
    pc->invMassB = bodyB->invMass;
    pc->localCenterA = bodyA->sweep.localCenter;
    pc->localCenterB = bodyB->sweep.localCenter;
    pc->invIA = bodyA->invI;
    pc->invIB = bodyB->invI;
    pc->localNormal = manifold->localNormal;
    pc->localPoint = manifold->localPoint;
    pc->pointCount = pointCount;
    pc->radiusA = radiusA;
    pc->radiusB = radiusB;
    pc->type = manifold->type;

    for (int j = 0; j < pointCount; ++j) {
      ManifoldPoint* cp = manifold->points + j;
      VelocityConstraintPoint* vcp = vc->points + j;
      if (warmStarting) {
        vcp->normalImpulse = dtRatio * cp->normalImpulse;
        vcp->tangentImpulse = dtRatio * cp->tangentImpulse;
      } else {
        vcp->normalImpulse = 0.0f;
        vcp->tangentImpulse = 0.0f;
      }
      vcp->rA.SetZero();
      vcp->rB.SetZero();
      vcp->normalMass = 0.0f;
      vcp->tangentMass = 0.0f;
      vcp->velocityBias = 0.0f;
      pc->localPoints[j] = cp->localPoint;
    }
  }
}

namespace {
struct WorldManifold {
  Vec2 normal;
  Vec2 points[kMaxManifoldPoints];
  void Initialize(const Manifold* manifold, const Xf& xfA, float radiusA, const Xf& xfB, float radiusB) {
    if (manifold->pointCount == 0) return;
    switch (manifold->type) {
      case kManifoldCircles: {
        normal.Set(1.0f, 0.0f);
        Vec2 pointA = Mul(xfA, manifold->localPoint);
        Vec2 pointB = Mul(xfB, manifold->points[0].localPoint);
        if (DistanceSquared(pointA, pointB) > kEpsilon * kEpsilon) {
          normal = pointB - pointA;
          normal.Normalize();
        }
        Vec2 cA = pointA + radiusA * normal;
        Vec2 cB = pointB - radiusB * normal;
        points[0] = 0.5f * (cA + cB);
      } break;
      case kManifoldFaceA: {
        normal = Mul(xfA.q, manifold->localNormal);
        Vec2 planePoint = Mul(xfA, manifold->localPoint);
        for (int i = 0; i < manifold->pointCount; ++i) {
          Vec2 clipPoint = Mul(xfB, manifold->points[i].localPoint);
          Vec2 cA = clipPoint + (radiusA - Dot(clipPoint - planePoint, normal)) * normal;
          Vec2 cB = clipPoint - radiusB * normal;
          points[i] = 0.5f * (cA + cB);
        }
      } break;
      case kManifoldFaceB: {
        normal = Mul(xfB.q, manifold->localNormal);
        Vec2 planePoint = Mul(xfB, manifold->localPoint);
        for (int i = 0; i < manifold->pointCount; ++i) {
          Vec2 clipPoint = Mul(xfA, manifold->points[i].localPoint);
          Vec2 cB = clipPoint + (radiusB - Dot(clipPoint - planePoint, normal)) * normal;
          Vec2 cA = clipPoint - radiusA * normal;
          points[i] = 0.5f * (cA + cB);
        }
        normal = -normal;
      } break;
    }
  }
};

struct PositionSolverManifold {
  Vec2 normal, point;
  float separation;
  void Initialize(const ContactPositionConstraint* pc, const Xf& xfA, const Xf& xfB, int index) {
    switch (pc->type) {
      case kManifoldCircles: {
        Vec2 pointA = Mul(xfA, pc->localPoint);
        Vec2 pointB = Mul(xfB, pc->localPoints[0]);
        normal = pointB - pointA;
        normal.Normalize();
        point = 0.5f * (pointA + pointB);
        separation = Dot(pointB - pointA, normal) - pc->radiusA - pc->radiusB;
      } break;
      case kManifoldFaceA: {
        normal = Mul(xfA.q, pc->localNormal);
        Vec2 planePoint = Mul(xfA, pc->localPoint);
        Vec2 clipPoint = Mul(xfB, pc->localPoints[index]);
        separation = Dot(clipPoint - planePoint, normal) - pc->radiusA - pc->radiusB;
        point = clipPoint;
      } break;
      case kManifoldFaceB: {
        normal = Mul(xfB.q, pc->localNormal);
        Vec2 planePoint = Mul(xfB, pc->localPoint);
        Vec2 clipPoint = Mul(xfA, pc->localPoints[index]);
        separation = Dot(clipPoint - planePoint, normal) - pc->radiusA - pc->radiusB;
        point = clipPoint;
        normal = -normal;
      } break;
    }
  }
};
}  // namespace

void ContactSolver::InitializeVelocityConstraints() {
  for (size_t i = 0; i < vcs.size(); ++i) {
    ContactVelocityConstraint* vc = &vcs[i];
    ContactPositionConstraint* pc = &pcs[i];
    float radiusA = pc->radiusA;
    float radiusB = pc->radiusB;
    Manifold* manifold = &contacts[vc->contactIndex]->manifold;
    int indexA = vc->indexA;
    int indexB = vc->indexB;
    float mA = vc->invMassA;
    float mB = vc->invMassB;
    float iA = vc->invIA;
    float iB = vc->invIB;
    Vec2 localCenterA = pc->localCenterA;
    Vec2 localCenterB = pc->localCenterB;
    Vec2 cA = positions[indexA].c;
    float aA = positions[indexA].a;
    Vec2 vA = velocities[indexA].v;
    float wA = velocities[indexA].w;
    Vec2 cB = positions[indexB].c;
    float aB = positions[indexB].a;
    Vec2 vB = velocities[indexB].v;
    float wB = velocities[indexB].w;

    Xf xfA, xfB;
    xfA.q.Set(aA);
    xfB.q.Set(aB);
    xfA.p = cA - Mul(xfA.q, localCenterA);
    xfB.p = cB - Mul(xfB.q, localCenterB);

    WorldManifold worldManifold;
    worldManifold.Initialize(manifold, xfA, radiusA, xfB, radiusB);
    vc->normal = worldManifold.normal;

    int pointCount = vc->pointCount;
    for (int j = 0; j < pointCount; ++j) {
      VelocityConstraintPoint* vcp = vc->points + j;
      vcp->rA = worldManifold.points[j] - cA;
      vcp->rB = worldManifold.points[j] - cB;
      float rnA = Cross(vcp->rA, vc->normal);
      float rnB = Cross(vcp->rB, vc->normal);
      float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
      vcp->normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
      Vec2 tangent = Cross(vc->normal, 1.0f);
      float rtA = Cross(vcp->rA, tangent);
      float rtB = Cross(vcp->rB, tangent);
      float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
      vcp->tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
      vcp->velocityBias = 0.0f;
      float vRel = Dot(vc->normal, vB + Cross(wB, vcp->rB) - vA - Cross(wA, vcp->rA));
      if (vRel < -kVelocityThreshold) vcp->velocityBias = -vc->restitution * vRel;
    }

    if (vc->pointCount == 2) {
      VelocityConstraintPoint* vcp1 = vc->points + 0;
      VelocityConstraintPoint* vcp2 = vc->points + 1;
      float rn1A = Cross(vcp1->rA, vc->normal);
      float rn1B = Cross(vcp1->rB, vc->normal);
      float rn2A = Cross(vcp2->rA, vc->normal);
      float rn2B = Cross(vcp2->rB, vc->normal);
      float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
      float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
      float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
      const float k_maxConditionNumber = 1000.0f;
      if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
        vc->K.ex.Set(k11, k12);
        vc->K.ey.Set(k12, k22);
        vc->normalMass = vc->K.GetInverse();
      } else {
        vc->pointCount = 1;
      }
    }
  }
}

void ContactSolver::WarmStart() {
  for (size_t i = 0; i < vcs.size(); ++i) {
    ContactVelocityConstraint* vc = &vcs[i];
    int indexA = vc->indexA;
    int indexB = vc->indexB;
    float mA = vc->invMassA;
    float iA = vc->invIA;
    float mB = vc->invMassB;
    float iB = vc->invIB;
    int pointCount = vc->pointCount;
    Vec2 vA = velocities[indexA].v;
    float wA = velocities[indexA].w;
    Vec2 vB = velocities[indexB].v;
    float wB = velocities[indexB].w;
    Vec2 normal = vc->normal;
    Vec2 tangent = Cross(normal, 1.0f);
    for (int j = 0; j < pointCount; ++j) {
      VelocityConstraintPoint* vcp = vc->points + j;
      Vec2 P = vcp->normalImpulse * normal + vcp->tangentImpulse * tangent;
      wA -= iA * Cross(vcp->rA, P);
      vA -= mA * P;
      wB += iB * Cross(vcp->rB, P);
      vB += mB * P;
    }
    velocities[indexA].v = vA;
    velocities[indexA].w = wA;
    velocities[indexB].v = vB;
    velocities[indexB].w = wB;
  }
}

void ContactSolver::SolveVelocityConstraints() {
  for (size_t i = 0; i < vcs.size(); ++i) {
    ContactVelocityConstraint* vc = &vcs[i];
    int indexA = vc->indexA;
    int indexB = vc->indexB;
    float mA = vc->invMassA;
    float iA = vc->invIA;
    float mB = vc->invMassB;
    float iB = vc->invIB;
    int pointCount = vc->pointCount;
    Vec2 vA = velocities[indexA].v;
    float wA = velocities[indexA].w;
    Vec2 vB = velocities[indexB].v;
    float wB = velocities[indexB].w;
    Vec2 normal = vc->normal;
    Vec2 tangent = Cross(normal, 1.0f);
    float friction = vc->friction;

    for (int j = 0; j < pointCount; ++j) {
      VelocityConstraintPoint* vcp = vc->points + j;
      Vec2 dv = vB + Cross(wB, vcp->rB) - vA - Cross(wA, vcp->rA);
      float vt = Dot(dv, tangent) - 0.0f;  // tangentSpeed = 0
      float lambda = vcp->tangentMass * (-vt);
      float maxFriction = friction * vcp->normalImpulse;
      float newImpulse = Clamp(vcp->tangentImpulse + lambda, -maxFriction, maxFriction);
      lambda = newImpulse - vcp->tangentImpulse;
      vcp->tangentImpulse = newImpulse;
      Vec2 P = lambda * tangent;
      vA -= mA * P;
      wA -= iA * Cross(vcp->rA, P);
      vB += mB * P;
      wB += iB * Cross(vcp->rB, P);
    }

    if (vc->pointCount == 1) {
      VelocityConstraintPoint* vcp = vc->points + 0;
      Vec2 dv = vB + Cross(wB, vcp->rB) - vA - Cross(wA, vcp->rA);
      float vn = Dot(dv, normal);
      float lambda = -vcp->normalMass * (vn - vcp->velocityBias);
      float newImpulse = Max(vcp->normalImpulse + lambda, 0.0f);
      lambda = newImpulse - vcp->normalImpulse;
      vcp->normalImpulse = newImpulse;
      Vec2 P = lambda * normal;
      vA -= mA * P;
      wA -= iA * Cross(vcp->rA, P);
      vB += mB * P;
      wB += iB * Cross(vcp->rB, P);
    } else {
      VelocityConstraintPoint* cp1 = vc->points + 0;
      VelocityConstraintPoint* cp2 = vc->points + 1;
      Vec2 a(cp1->normalImpulse, cp2->normalImpulse);
      Vec2 dv1 = vB + Cross(wB, cp1->rB) - vA - Cross(wA, cp1->rA);
      Vec2 dv2 = vB + Cross(wB, cp2->rB) - vA - Cross(wA, cp2->rA);
      float vn1 = Dot(dv1, normal);
      float vn2 = Dot(dv2, normal);
      Vec2 b;
      b.x = vn1 - cp1->velocityBias;
      b.y = vn2 - cp2->velocityBias;
      b -= Mul(vc->K, a);
      for (;;) {
        Vec2 x = -Mul(vc->normalMass, b);
        if (x.x >= 0.0f && x.y >= 0.0f) {
          Vec2 d = x - a;
          Vec2 P1 = d.x * normal;
          Vec2 P2 = d.y * normal;
          vA -= mA * (P1 + P2);
          wA -= iA * (Cross(cp1->rA, P1) + Cross(cp2->rA, P2));
          vB += mB * (P1 + P2);
          wB += iB * (Cross(cp1->rB, P1) + Cross(cp2->rB, P2));
          cp1->normalImpulse = x.x;
          cp2->normalImpulse = x.y;
          break;
        }
        x.x = -cp1->normalMass * b.x;
        x.y = 0.0f;
        vn1 = 0.0f;
        vn2 = vc->K.ex.y * x.x + b.y;
        if (x.x >= 0.0f && vn2 >= 0.0f) {
          Vec2 d = x - a;
          Vec2 P1 = d.x * normal;
          Vec2 P2 = d.y * normal;
          vA -= mA * (P1 + P2);
          wA -= iA * (Cross(cp1->rA, P1) + Cross(cp2->rA, P2));
          vB += mB * (P1 + P2);
          wB += iB * (Cross(cp1->rB, P1) + Cross(cp2->rB, P2));
          cp1->normalImpulse = x.x;
          cp2->normalImpulse = x.y;
          break;
        }
        x.x = 0.0f;
        x.y = -cp2->normalMass * b.y;
        vn1 = vc->K.ey.x * x.y + b.x;
        vn2 = 0.0f;
        if (x.y >= 0.0f && vn1 >= 0.0f) {
          Vec2 d = x - a;
          Vec2 P1 = d.x * normal;
          Vec2 P2 = d.y * normal;
          vA -= mA * (P1 + P2);
          wA -= iA * (Cross(cp1->rA, P1) + Cross(cp2->rA, P2));
          vB += mB * (P1 + P2);
          wB += iB * (Cross(cp1->rB, P1) + Cross(cp2->rB, P2));
          cp1->normalImpulse = x.x;
          cp2->normalImpulse = x.y;
          break;
        }
        x.x = 0.0f;
        x.y = 0.0f;
        vn1 = b.x;
        vn2 = b.y;
        if (vn1 >= 0.0f && vn2 >= 0.0f) {
          Vec2 d = x - a;
          Vec2 P1 = d.x * normal;
          Vec2 P2 = d.y * normal;
          vA -= mA * (P1 + P2);
          wA -= iA * (Cross(cp1->rA, P1) + Cross(cp2->rA, P2));
          vB += mB * (P1 + P2);
          wB += iB * (Cross(cp1->rB, P1) + Cross(cp2->rB, P2));
          cp1->normalImpulse = x.x;
          cp2->normalImpulse = x.y;
          break;
        }
        break;
      }
    }
    velocities[indexA].v = vA;
    velocities[indexA].w = wA;
    velocities[indexB].v = vB;
    velocities[indexB].w = wB;
  }
}

void ContactSolver::StoreImpulses() {
  for (size_t i = 0; i < vcs.size(); ++i) {
    ContactVelocityConstraint* vc = &vcs[i];
    Manifold* manifold = &contacts[vc->contactIndex]->manifold;
    for (int j = 0; j < vc->pointCount; ++j) {
      manifold->points[j].normalImpulse = vc->points[j].normalImpulse;
      manifold->points[j].tangentImpulse = vc->points[j].tangentImpulse;
    }
  }
}

bool ContactSolver::SolvePositionConstraints() {
  float minSeparation = 0.0f;
  for (size_t i = 0; i < pcs.size(); ++i) {
    ContactPositionConstraint* pc = &pcs[i];
    int indexA = pc->indexA;
    int indexB = pc->indexB;
    Vec2 localCenterA = pc->localCenterA;
    float mA = pc->invMassA;
    float iA = pc->invIA;
    Vec2 localCenterB = pc->localCenterB;
    float mB = pc->invMassB;
    float iB = pc->invIB;
    int pointCount = pc->pointCount;
    Vec2 cA = positions[indexA].c;
    float aA = positions[indexA].a;
    Vec2 cB = positions[indexB].c;
    float aB = positions[indexB].a;
    for (int j = 0; j < pointCount; ++j) {
      Xf xfA, xfB;
      xfA.q.Set(aA);
      xfB.q.Set(aB);
      xfA.p = cA - Mul(xfA.q, localCenterA);
      xfB.p = cB - Mul(xfB.q, localCenterB);
      PositionSolverManifold psm;
      psm.Initialize(pc, xfA, xfB, j);
      Vec2 normal = psm.normal;
      Vec2 point = psm.point;
      float separation = psm.separation;
      Vec2 rA = point - cA;
      Vec2 rB = point - cB;
      minSeparation = Min(minSeparation, separation);
      float C = Clamp(kBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
      float rnA = Cross(rA, normal);
      float rnB = Cross(rB, normal);
      float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
      float impulse = K > 0.0f ? -C / K : 0.0f;
      Vec2 P = impulse * normal;
      cA -= mA * P;
      aA -= iA * Cross(rA, P);
      cB += mB * P;
      aB += iB * Cross(rB, P);
    }
    positions[indexA].c = cA;
    positions[indexA].a = aA;
    positions[indexB].c = cB;
    positions[indexB].a = aB;
  }
  return minSeparation >= -3.0f * kLinearSlop;
}

bool ContactSolver::SolveTOIPositionConstraints(int toiIndexA, int toiIndexB) {
  float minSeparation = 0.0f;
  for (size_t i = 0; i < pcs.size(); ++i) {
    ContactPositionConstraint* pc = &pcs[i];
    int indexA = pc->indexA;
    int indexB = pc->indexB;
    Vec2 localCenterA = pc->localCenterA;
    Vec2 localCenterB = pc->localCenterB;
    int pointCount = pc->pointCount;
    float mA = 0.0f;
    float iA = 0.0f;
    if (indexA == toiIndexA || indexA == toiIndexB) {
      mA = pc->invMassA;
      iA = pc->invIA;
    }
    float mB = 0.0f;
    float iB = 0.0f;
    if (indexB == toiIndexA || indexB == toiIndexB) {
      mB = pc->invMassB;
      iB = pc->invIB;
    }
    Vec2 cA = positions[indexA].c;
    float aA = positions[indexA].a;
    Vec2 cB = positions[indexB].c;
    float aB = positions[indexB].a;
    for (int j = 0; j < pointCount; ++j) {
      Xf xfA, xfB;
      xfA.q.Set(aA);
      xfB.q.Set(aB);
      xfA.p = cA - Mul(xfA.q, localCenterA);
      xfB.p = cB - Mul(xfB.q, localCenterB);
      PositionSolverManifold psm;
      psm.Initialize(pc, xfA, xfB, j);
      Vec2 normal = psm.normal;
      Vec2 point = psm.point;
      float separation = psm.separation;
      Vec2 rA = point - cA;
      Vec2 rB = point - cB;
      minSeparation = Min(minSeparation, separation);
      float C = Clamp(kToiBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
      float rnA = Cross(rA, normal);
      float rnB = Cross(rB, normal);
      float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
      float impulse = K > 0.0f ? -C / K : 0.0f;
      Vec2 P = impulse * normal;
      cA -= mA * P;
      aA -= iA * Cross(rA, P);
      cB += mB * P;
      aB += iB * Cross(rB, P);
    }
    positions[indexA].c = cA;
    positions[indexA].a = aA;
    positions[indexB].c = cB;
    positions[indexB].a = aB;
  }
  return minSeparation >= -1.5f * kLinearSlop;
}

}  // namespace kbo
