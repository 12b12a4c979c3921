// kb_step.cuh -- one lane group per environment; the whole KilobotsEnv.step
// (gym_kilobots/envs/kilobots_env.py:161-215) for that environment runs inside one kernel launch with
// bodies, fat AABBs, contact ids/flags, controllers and the solver's constraint pool resident in
// shared memory across all sub-steps of the action.
//
// Phases per sub-step (reference call sites in brackets):
//   light_step      [lib/light.py:59-75, 300-316, 237-253; kilobots_env.py:171-172]
//   sense_control   [lib/kilobot.py:54-55,188-189; lib/light.py:176-189; lib/kilobot.py:86-127,191-203,253-258,294-300,318-333]
//   collide         [Box2D b2ContactManager::Collide + b2Contact::Update, via b2World.Step kilobots_env.py:187]
//   solve           [b2World::Solve: island DFS, b2Island::Solve, b2ContactSolver; sleeping; SynchronizeFixtures]
//   find_new        [b2BroadPhase::UpdatePairs + b2ContactManager::AddPair]
//   solve_toi       [b2World::SolveTOI, kb_toi.cuh]
// Ordering contract: the persistent contact array is kept in creation order, so Box2D's LIFO world
// list / per-body contact lists are "descending array index"; the island DFS, the constraint order
// and therefore every float32 result follow Box2D's sequential-impulse order exactly.  Within that
// order, constraints are executed level by level (a level = constraints whose dynamic bodies are
// disjoint and whose predecessors are done), which is bit-identical to the sequential sweep.
#pragma once
#include "kb_narrow.cuh"

namespace kb {

struct KernelArgs {
  Layout L;
  float* blobs;                 // [E][blobWords]
  const int32_t* envScene;      // [E]
  const ProxyConst* proxies;    // [S][Pp]
  const BodyConst* bodies;      // [S][Bp]
  const SceneConst* scenes;     // [S]
  const LightConst* lights;     // [numLights]
  int32_t numEnvs;
  // step
  const double* action;
  int32_t actionMode;
  float* obsKilobots;
  float* obsObjects;
  double* obsLight;
  float* reward;
  uint8_t* done;
  int32_t* status;
  // reset
  const uint8_t* mask;
  const double* pose;
  const double* lightInit;
  const double* kbVel;
};

// pool field indices (velocity phase)
#define PF_IDX 0   /* bA | bB<<8 | pointCount<<16 | general<<20 | type<<24 */
#define PF_AUX 1   /* contact index | island<<16 */
#define PF_NX 2
#define PF_NY 3
#define PF_RAX 4
#define PF_RAY 5
#define PF_RBX 6
#define PF_RBY 7
#define PF_NMASS 8
#define PF_NIMP 9
#define POOL_FIELDS 10

template <int LPE>
struct Sim {
  Group<LPE> g;
  const Layout& L;
  uint32_t* sm;                 // this env's shared memory (words)
  float* blob;                  // this env's state image in HBM
  const ProxyConst* px;         // scene proxies
  const BodyConst* bc;          // scene bodies
  const LightConst* lights;
  int S;                        // index of the static table in the body arrays (== L.B)

  __device__ __forceinline__ Sim(const Layout& l) : L(l) {}

  // ---- typed views into shared memory
  __device__ __forceinline__ float4& pos4(int b) { return reinterpret_cast<float4*>(sm + L.oPos)[b]; }   // cx cy a sleepTime
  __device__ __forceinline__ float4& vel4(int b) { return reinterpret_cast<float4*>(sm + L.oVel)[b]; }   // vx vy w flags
  __device__ __forceinline__ float4& xf4(int b) { return reinterpret_cast<float4*>(sm + L.oXf)[b]; }     // px py qs qc
  __device__ __forceinline__ float4& fat4(int p) { return reinterpret_cast<float4*>(sm + L.oFat)[p]; }   // lx ly ux uy
  __device__ __forceinline__ float4& sweep4(int b) { return reinterpret_cast<float4*>(sm + L.sSweep)[b]; } // c0x c0y a0 alpha0
  __device__ __forceinline__ float4& bc4(int b) { return reinterpret_cast<float4*>(sm + L.sBc)[b]; }     // invMass invI lcx lcy
  __device__ __forceinline__ uint32_t& cpair(int i) { return sm[L.oPair + i]; }
  __device__ __forceinline__ uint32_t& cinfo(int i) { return sm[L.oInfo + i]; }
  __device__ __forceinline__ uint32_t& hdr(int i) { return sm[L.oHdr + i]; }
  __device__ __forceinline__ int32_t& isl(int b) { return reinterpret_cast<int32_t*>(sm + L.sIsl)[b]; }
  __device__ __forceinline__ uint32_t& islflag(int i) { return sm[L.sIslMin + i]; }
  __device__ __forceinline__ uint32_t& misc(int i) { return sm[L.sMisc + i]; }
  __device__ __forceinline__ uint32_t& adj(int p, int hi) { return sm[L.sAdj + 2 * p + hi]; }
  __device__ __forceinline__ float& pool(int f, int q) { return reinterpret_cast<float*>(sm + L.sPool)[f * L.Kmax + q]; }
  __device__ __forceinline__ uint32_t& poolu(int f, int q) { return sm[L.sPool + f * L.Kmax + q]; }
  // general contacts occupy 3 consecutive slots: word w of the 30-word record
  __device__ __forceinline__ float& gw(int q, int w) { return pool(w % POOL_FIELDS, q + w / POOL_FIELDS); }
  __device__ __forceinline__ double* lightState() { return reinterpret_cast<double*>(sm + L.oLight); }
  __device__ __forceinline__ double* ctrl(int k) { return reinterpret_cast<double*>(sm + L.oCtrl) + 4 * k; }
  __device__ __forceinline__ unsigned long long* counters() { return reinterpret_cast<unsigned long long*>(sm + L.oCnt); }
  __device__ __forceinline__ float* manifoldRec(int i) { return blob + L.oMan + MR_WORDS * i; }

  __device__ __forceinline__ Xf bodyXf(int b) {
    float4 x = xf4(b);
    Xf t;
    t.p = mk(x.x, x.y);
    t.q.s = x.z;
    t.q.c = x.w;
    return t;
  }
  __device__ __forceinline__ bool awake(int b) { return (f2u(vel4(b).w) & BF_AWAKE) != 0u; }
  // b2Body::SetAwake(true): only acts on sleeping bodies
  __device__ __forceinline__ void wake(int b) {
    if (b == S) return;
    uint32_t f = f2u(vel4(b).w);
    if ((f & BF_AWAKE) == 0u) {
      reinterpret_cast<float*>(&vel4(b))[3] = u2f(f | BF_AWAKE);
      reinterpret_cast<float*>(&pos4(b))[3] = 0.0f;
    }
  }

  // ------------------------------------------------------------------------- state I/O
  __device__ void loadState() {
    const float4* src = reinterpret_cast<const float4*>(blob);
    float4* dst = reinterpret_cast<float4*>(sm);
    const int n4 = L.stateWords >> 2;
    for (int i = g.lane; i < n4; i += LPE) dst[i] = src[i];
    g.sync();
  }
  __device__ void storeState() {
    g.sync();
    float4* dst = reinterpret_cast<float4*>(blob);
    const float4* src = reinterpret_cast<const float4*>(sm);
    const int n4 = L.stateWords >> 2;
    for (int i = g.lane; i < n4; i += LPE) dst[i] = src[i];
  }
  // scratch that is constant for the launch: body constants, static-table slot, adjacency masks
  __device__ void initScratch() {
    for (int b = g.lane; b <= L.B; b += LPE) {
      if (b < L.B) {
        const BodyConst* c = bc + b;
        bc4(b) = make_float4(__ldg(&c->invMass), __ldg(&c->invI), __ldg(&c->lcx), __ldg(&c->lcy));
      } else {
        bc4(b) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        pos4(b) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        vel4(b) = make_float4(0.0f, 0.0f, 0.0f, u2f(BF_AWAKE));
        xf4(b) = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
        sweep4(b) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
    }
    for (int p = g.lane; p < 2 * L.P; p += LPE) sm[L.sAdj + p] = 0u;
    g.sync();
    const int nC = (int)hdr(H_NC);
    for (int i = g.lane; i < nC; i += LPE) {
      uint32_t pr = cpair(i);
      int pa = pr & 0xFFFF, pb = pr >> 16;
      atomicOr(&adj(pa, pb >> 5), 1u << (pb & 31));
      atomicOr(&adj(pb, pa >> 5), 1u << (pa & 31));
    }
    g.sync();
  }

  // --------------------------------------------------------------------------- lights
  __device__ __forceinline__ static double clipd(double a, double lo, double hi) {
    double m = a > lo ? a : lo;
    return m < hi ? m : hi;
  }
  __device__ void lightStep(const double* action) {
    if (g.lane == 0) {
      double* ls = lightState();
      int so = 0, ao = 0;
      for (int l = 0; l < L.numLights; ++l) {
        const LightConst& lc = lights[l];
        if (lc.type == KB_LIGHT_CIRCULAR) {
          double a0 = clipd(action[ao], lc.alo[0], lc.ahi[0]);
          double a1 = clipd(action[ao + 1], lc.alo[1], lc.ahi[1]);
          const double dt = 1. / 10;
          if (lc.relative) {
            ls[so] += a0 * dt;
            ls[so + 1] += a1 * dt;
          } else {
            ls[so] = a0;
            ls[so + 1] = a1;
          }
          ls[so] = clipd(ls[so], lc.blo[0], lc.bhi[0]);
          ls[so + 1] = clipd(ls[so + 1], lc.blo[1], lc.bhi[1]);
          so += 2;
          ao += 2;
        } else if (lc.type == KB_LIGHT_MOMENTUM) {
          const double dt = 1. / 10;
          double a0 = clipd(action[ao], lc.alo[0], lc.ahi[0]);
          double a1 = clipd(action[ao + 1], lc.alo[1], lc.ahi[1]);
          ls[so + 2] += a0 * dt;
          ls[so + 3] += a1 * dt;
          double n = sqrt(ls[so + 2] * ls[so + 2] + ls[so + 3] * ls[so + 3]);
          if (n > lc.maxVel) {
            double f = lc.maxVel / n;
            ls[so + 2] *= f;
            ls[so + 3] *= f;
          }
          ls[so] += ls[so + 2] * dt;
          ls[so + 1] += ls[so + 3] * dt;
          ls[so] = clipd(ls[so], lc.blo[0], lc.bhi[0]);
          ls[so + 1] = clipd(ls[so + 1], lc.blo[1], lc.bhi[1]);
          so += 4;
          ao += 2;
        } else {
          double a = clipd(action[ao], lc.alo[0], lc.ahi[0]);
          const double pi = 3.141592653589793;
          if (a < -pi) a += 2 * pi;
          if (a > pi) a -= 2 * pi;
          ls[so] = a;
          so += 1;
          ao += 1;
        }
      }
    }
    g.sync();
  }

  __device__ __forceinline__ void lightValueGrad(const LightConst& lc, const double* ls, double sx, double sy,
                                                 double* value, double* gx, double* gy) {
    if (lc.type == KB_LIGHT_LINEAR) {
      double vx, vy;
      kb_sincosd(ls[0], &vy, &vx);
      *value = vx * sx + vy * sy;
      *gx = vx;
      *gy = vy;
      return;
    }
    double g0 = -1 * (sx - ls[0]);
    double g1 = -1 * (sy - ls[1]);
    double norm = sqrt(g0 * g0 + g1 * g1);
    double v = 1.0;
    v -= norm / lc.radius;
    v = v < 1. ? v : 1.;
    v = v > .0 ? v : .0;
    v *= 255;
    if (norm == 0.0) {
      g0 = 0.0;
      g1 = 0.0;
    } else {
      g0 /= norm;
      g1 /= norm;
    }
    if (norm > lc.radius) {
      g0 *= .0;
      g1 *= .0;
    }
    *value = v;
    *gx = g0;
    *gy = g1;
  }

  // b2Body::SetLinearVelocity / SetAngularVelocity on a kilobot from its own lane
  __device__ __forceinline__ void setLinearVelocity(int b, V2 v) {
    if (dot(v, v) > 0.0f) wake(b);
    float* vv = reinterpret_cast<float*>(&vel4(b));
    vv[0] = v.x;
    vv[1] = v.y;
  }
  __device__ __forceinline__ void setAngularVelocity(int b, float w) {
    if (w * w > 0.0f) wake(b);
    reinterpret_cast<float*>(&vel4(b))[2] = w;
  }

  __device__ void setKilobotActions(const double* action) {
    const double hpi = 0.5 * 3.141592653589793;
    const double pi = 3.141592653589793;
    for (int k = g.lane; k < L.N; k += LPE) {
      const int kind = __ldg(&bc[L.M + k].kind);
      double* c = ctrl(k);
      const double* a = action ? action + 2 * k : nullptr;
      if (kind == KB_KILOBOT_VELOCITY) {
        if (a) {
          c[0] = clipd(a[0], .0, 0.01);
          c[1] = clipd(a[1], -hpi, hpi);
        } else {
          c[0] = .0;
          c[1] = .0;
        }
      } else if (kind == KB_KILOBOT_ACCELERATION) {
        if (a) {
          c[2] = clipd(a[0], -.005, .005);
          c[3] = clipd(a[1], -.2 * pi, .2 * pi);
        } else {
          c[2] = .0;
          c[3] = .0;
        }
      }
    }
    g.sync();
  }

  __device__ void senseControl() {
    const double* ls = lightState();
    for (int k = g.lane; k < L.N; k += LPE) {
      const int b = L.M + k;
      const int kind = __ldg(&bc[b].kind);
      const Xf xf = bodyXf(b);
      double value = 0.0, gx = 0.0, gy = 0.0;
      if (L.numLights > 0) {
        double sx, sy;
        if (kind == KB_KILOBOT_SIMPLE_PHOTOTAXIS) {
          sx = (double)xf.p.x / 25.0;
          sy = (double)xf.p.y / 25.0;
        } else {
          V2 lp = mk((float)(25.0 * 0.0), (float)(25.0 * -0.0165));
          V2 wp = xmul(xf, lp);
          sx = (double)wp.x / 25.0;
          sy = (double)wp.y / 25.0;
        }
        if (L.numLights == 1) {
          lightValueGrad(lights[0], ls, sx, sy, &value, &gx, &gy);
        } else {
          double best = 0.0;
          int so = 0;
          for (int l = 0; l < L.numLights; ++l) {
            double v, g0, g1;
            lightValueGrad(lights[l], ls + so, sx, sy, &v, &g0, &g1);
            value += v;
            if (l == 0 || v > best) {
              best = v;
              gx = g0;
              gy = g1;
            }
            so += lights[l].type == KB_LIGHT_MOMENTUM ? 4 : (lights[l].type == KB_LIGHT_LINEAR ? 1 : 2);
          }
        }
      }
      double* c = ctrl(k);
      switch (kind) {
        case KB_KILOBOT_PHOTOTAXIS: {
          // c[0] threshold, c[1] turnRight, c[2] updateCounter, c[3] noChangeCounter
          int upd = (int)c[2];
          int turnRight = (int)c[1];
          if (upd % 6) {
            upd += 1;
          } else {
            upd += 1;
            int noChange = (int)c[3];
            if (value > c[0] || noChange >= 15) {
              c[0] = value + .01;
              turnRight = turnRight ? 0 : 1;
              noChange = 0;
            } else {
              noChange += 1;
            }
            c[3] = (double)noChange;
            c[1] = (double)turnRight;
          }
          c[2] = (double)upd;
          const float tx = turnRight ? L.transRight[0] : L.transLeft[0];
          const float ty = turnRight ? L.transRight[1] : L.transLeft[1];
          const float w = turnRight ? L.omegaRight : L.omegaLeft;
          V2 wv = rmul(xf.q, mk(tx, ty));
          V2 lv = mk(wv.x / 25.0f, wv.y / 25.0f);
          lv = mk(lv.x / L.dt, lv.y / L.dt);
          lv = mk(lv.x * 25.0f, lv.y * 25.0f);
          setAngularVelocity(b, w);
          setLinearVelocity(b, lv);
        } break;
        case KB_KILOBOT_SIMPLE_PHOTOTAXIS: {
          double mx = gx, my = gy;
          double n = sqrt(mx * mx + my * my);
          if (n > 0.01) {
            mx = mx / n * 0.01;
            my = my / n * 0.01;
          }
          mx *= 25.0;
          my *= 25.0;
          setLinearVelocity(b, mk((float)mx, (float)my));
          // linearDamping = 0: kept as a per-kind constant, see integrateVelocities()
        } break;
        case KB_KILOBOT_ACCELERATION: {
          const double hpi = 0.5 * 3.141592653589793;
          const double dt = 1. / 10;
          c[0] += c[2] * dt;
          c[1] += c[3] * dt;
          c[0] = c[0] > .0 ? c[0] : .0;
          c[1] = c[1] > -hpi ? c[1] : -hpi;
          c[0] = c[0] < 0.01 ? c[0] : 0.01;
          c[1] = c[1] < hpi ? c[1] : hpi;
        }  // fallthrough
        case KB_KILOBOT_VELOCITY: {
          double ang = (double)pos4(b).z;
          double lx, ly;
          kb_sincosd(ang, &ly, &lx);
          lx *= c[0] * 25.0;
          ly *= c[0] * 25.0;
          setLinearVelocity(b, mk((float)lx, (float)ly));
          setAngularVelocity(b, (float)c[1]);
        } break;
        default: break;
      }
    }
    g.sync();
  }

  // --------------------------------------------------------------------------- narrowphase
  __device__ __forceinline__ void evaluate(Manifold& m, int pa, int pb, int bA, int bB) {
    const ProxyConst* A = px + pa;
    const ProxyConst* Bp = px + pb;
    const int tA = __ldg(&A->type), tB = __ldg(&Bp->type);
    const Xf xfA = bodyXf(bA), xfB = bodyXf(bB);
    if (tA == SHAPE_CIRCLE) {
      collide_circles(m, __ldg(&A->radius), xfA, __ldg(&Bp->radius), xfB);
    } else if (tA == SHAPE_POLYGON) {
      if (tB == SHAPE_CIRCLE) collide_polygon_circle(m, A, xfA, __ldg(&Bp->radius), xfB);
      else collide_polygons(m, A, xfA, Bp, xfB);
    } else {
      if (tB == SHAPE_CIRCLE) collide_edge_circle(m, A, xfA, __ldg(&Bp->radius), xfB);
      else collide_edge_polygon(m, A, xfA, Bp, xfB);
    }
  }

  // b2Contact::Update for contact i (this lane).  Returns true if touching changed.
  __device__ __forceinline__ bool updateContact(int i) {
    const uint32_t pr = cpair(i);
    const int pa = pr & 0xFFFF, pb = pr >> 16;
    const int bA = __ldg(&px[pa].body), bB = __ldg(&px[pb].body);
    uint32_t info = cinfo(i);
    const int oldPC = (info & CI_PC_MASK) >> CI_PC_SHIFT;
    const bool wasTouching = (info & CI_TOUCHING) != 0u;
    Manifold m;
    m.pointCount = 0;
    m.type = 0;
    m.lnx = m.lny = m.lpx = m.lpy = 0.0f;
    m.px[0] = m.py[0] = m.px[1] = m.py[1] = 0.0f;
    m.id[0] = m.id[1] = 0u;
    evaluate(m, pa, pb, bA, bB);
    const bool touching = m.pointCount > 0;
    float* rec = manifoldRec(i);
    if (touching) {
      m.ni[0] = m.ni[1] = m.ti[0] = m.ti[1] = 0.0f;
      if (oldPC > 0) {
        const float4 r1 = reinterpret_cast<const float4*>(rec)[1];  // p0x p0y p0n p0t
        const float4 r2 = reinterpret_cast<const float4*>(rec)[2];  // p0id p1x p1y p1n
        const float4 r3 = reinterpret_cast<const float4*>(rec)[3];  // p1t p1id type pad
        const uint32_t oid0 = f2u(r2.x), oid1 = f2u(r3.y);
        for (int k = 0; k < m.pointCount; ++k) {
          const uint32_t id2 = k == 0 ? m.id[0] : m.id[1];
          float ni = 0.0f, ti = 0.0f;
          if (oid0 == id2) {
            ni = r1.z;
            ti = r1.w;
          } else if (oldPC > 1 && oid1 == id2) {
            ni = r2.w;
            ti = r3.x;
          }
          if (k == 0) { m.ni[0] = ni; m.ti[0] = ti; }
          else { m.ni[1] = ni; m.ti[1] = ti; }
        }
      }
      float4* o = reinterpret_cast<float4*>(rec);
      o[0] = make_float4(m.lnx, m.lny, m.lpx, m.lpy);
      o[1] = make_float4(m.px[0], m.py[0], m.ni[0], m.ti[0]);
      o[2] = make_float4(u2f(m.id[0]), m.px[1], m.py[1], m.ni[1]);
      o[3] = make_float4(m.ti[1], u2f(m.id[1]), u2f((uint32_t)m.type | ((uint32_t)m.pointCount << 8)), 0.0f);
    }
    info |= CI_ENABLED;
    info = touching ? (info | CI_TOUCHING) : (info & ~CI_TOUCHING);
    info = (info & ~CI_PC_MASK) | ((uint32_t)m.pointCount << CI_PC_SHIFT);
    cinfo(i) = info;
    return touching != wasTouching;
  }

  // b2ContactManager::Collide.  World-list order is descending array index; a sleeping pair is
  // only visited if an earlier (higher index) contact woke one of its bodies.
  __device__ void collide() {
    const int nC = (int)hdr(H_NC);
    if (nC == 0) return;
    const uint32_t CI_DESTROY = 1u << 31, CI_DONE = 1u << 30;
    // any sleeping dynamic body?
    bool sleepy = false;
    for (int b = g.lane; b < L.B; b += LPE) sleepy |= !awake(b);
    const bool anyAsleep = g.any(sleepy);
    int32_t* wakeAt = reinterpret_cast<int32_t*>(sm + L.sStack);  // reuse: per body, highest waking contact index
    if (anyAsleep) {
      for (int b = g.lane; b <= L.B; b += LPE) wakeAt[b] = awake(b) ? 0x7FFFFFFF : -1;
      g.sync();
    }
    bool anyDestroyed = false;
    for (int pass = 0;; ++pass) {
      bool progressed = false;
      for (int base = 0; base < nC; base += LPE) {
        const int i = base + g.lane;
        bool doit = false;
        int pa = 0, pb = 0, bA = S, bB = S;
        if (i < nC) {
          const uint32_t info = cinfo(i);
          if ((info & CI_DONE) == 0u) {
            const uint32_t pr = cpair(i);
            pa = pr & 0xFFFF;
            pb = pr >> 16;
            bA = __ldg(&px[pa].body);
            bB = __ldg(&px[pb].body);
            if (!anyAsleep) {
              doit = true;
            } else {
              const bool actA = bA != S && wakeAt[bA] > i;
              const bool actB = bB != S && wakeAt[bB] > i;
              doit = actA || actB;
            }
          }
        }
        bool wakeEvent = false;
        if (doit) {
          const float4 fa = fat4(pa), fb = fat4(pb);
          // b2TestOverlap
          const bool overlap = !(fb.x - fa.z > 0.0f || fb.y - fa.w > 0.0f || fa.x - fb.z > 0.0f || fa.y - fb.w > 0.0f);
          uint32_t info = cinfo(i);
          if (!overlap) {
            wakeEvent = (info & CI_PC_MASK) != 0u;
            cinfo(i) = info | CI_DESTROY | CI_DONE;
            atomicAnd(&adj(pa, pb >> 5), ~(1u << (pb & 31)));
            atomicAnd(&adj(pb, pa >> 5), ~(1u << (pa & 31)));
          } else {
            wakeEvent = updateContact(i);
            cinfo(i) |= CI_DONE;
          }
          if (wakeEvent && anyAsleep) {
            if (bA != S) atomicMax(&wakeAt[bA], i);
            if (bB != S) atomicMax(&wakeAt[bB], i);
          }
        }
        progressed |= doit;
        anyDestroyed |= doit && (cinfo(i) & CI_DESTROY) != 0u;
        g.sync();
      }
      if (!anyAsleep) break;
      if (!g.any(progressed)) break;
    }
    if (anyAsleep) {
      g.sync();
      for (int b = g.lane; b < L.B; b += LPE)
        if (wakeAt[b] >= 0) wake(b);
      g.sync();
    }
    // clear DONE marks; stable compaction if anything was destroyed
    const bool compact = g.any(anyDestroyed);
    int out = 0;
    for (int base = 0; base < nC; base += LPE) {
      const int i = base + g.lane;
      uint32_t info = 0u, pr = 0u;
      bool keep = false;
      if (i < nC) {
        info = cinfo(i);
        pr = cpair(i);
        keep = (info & CI_DESTROY) == 0u;
        info &= ~(CI_DONE | CI_DESTROY);
      }
      if (!compact) {
        if (i < nC) cinfo(i) = info;
        continue;
      }
      const uint32_t m = g.ballot(keep);
      const int dst = out + __popc(m & g.lt());
      float4 r0, r1, r2, r3;
      const bool moveRec = keep && dst != i && (info & CI_PC_MASK) != 0u;
      if (moveRec) {
        const float4* rec = reinterpret_cast<const float4*>(manifoldRec(i));
        r0 = rec[0]; r1 = rec[1]; r2 = rec[2]; r3 = rec[3];
      }
      g.sync();
      if (keep) {
        cinfo(dst) = info;
        cpair(dst) = pr;
        if (moveRec) {
          float4* rec = reinterpret_cast<float4*>(manifoldRec(dst));
          rec[0] = r0; rec[1] = r1; rec[2] = r2; rec[3] = r3;
        }
      }
      out += __popc(m);
      g.sync();
    }
    if (compact) {
      if (g.lane == 0) hdr(H_NC) = (uint32_t)out;
      g.sync();
    }
  }

  // ------------------------------------------------------------------------------ solver
  // schedule entry e -> (order position p, pool slot q)
  __device__ __forceinline__ uint32_t& entry(int e) { return sm[L.sEslot + e]; }
  __device__ __forceinline__ uint32_t& ordC(int p) { return sm[L.sOrder + p]; }            // contact idx | island << 16
  __device__ __forceinline__ uint32_t& ordB(int p) { return sm[L.sOrder + L.Kmax + p]; }   // bA | bB << 8 | size << 16
  __device__ __forceinline__ uint32_t& lvlOff(int l) { return sm[L.sLvlOff + l]; }

  struct VelBody {
    V2 v;
    float w;
  };

  // b2ContactSolver ctor + InitializeVelocityConstraints for schedule entry e.  fresh == true is the
  // TOI island variant: rotations rebuilt from the (corrected) angles, no warm starting.
  __device__ __forceinline__ void initConstraint(int e, bool fresh = false) {
    const uint32_t en = entry(e);
    const int p = en & 0xFFFF, q = en >> 16;
    const uint32_t oc = ordC(p), ob = ordB(p);
    const int ci = oc & 0xFFFF;
    const int bA = ob & 0xFF, bB = (ob >> 8) & 0xFF;
    const bool general = ((ob >> 16) & 0xFF) > 1;
    const uint32_t pr = cpair(ci);
    const int pa = pr & 0xFFFF, pb = pr >> 16;
    const float4* rec = reinterpret_cast<const float4*>(manifoldRec(ci));
    const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3];
    const uint32_t tp = f2u(r3.z);
    const int type = tp & 0xFF, pointCount = (tp >> 8) & 0xFF;
    const float radiusA = __ldg(&px[pa].radius), radiusB = __ldg(&px[pb].radius);
    const float4 cA4 = pos4(bA), cB4 = pos4(bB);
    const float4 vA4 = vel4(bA), vB4 = vel4(bB);
    const float4 kA = bc4(bA), kB = bc4(bB);
    const float mA = kA.x, iA = kA.y, mB = kB.x, iB = kB.y;
    const V2 cA = mk(cA4.x, cA4.y), cB = mk(cB4.x, cB4.y);
    const V2 vA = mk(vA4.x, vA4.y), vB = mk(vB4.x, vB4.y);
    const float wA = vA4.z, wB = vB4.z;
    // xf from (c, a): q == the body's current xf.q (b2Rot::Set(sweep.a) is what produced it)
    Xf xfA, xfB;
    {
      if (fresh) {
        xfA.q = rot_set(cA4.z);
        xfB.q = rot_set(cB4.z);
      } else {
        const float4 xa = xf4(bA), xb = xf4(bB);
        xfA.q.s = xa.z; xfA.q.c = xa.w;
        xfB.q.s = xb.z; xfB.q.c = xb.w;
      }
      xfA.p = cA - rmul(xfA.q, mk(kA.z, kA.w));
      xfB.p = cB - rmul(xfB.q, mk(kB.z, kB.w));
    }
    // b2WorldManifold::Initialize
    V2 normal, pts[2];
    const V2 lp = mk(r0.z, r0.w), ln = mk(r0.x, r0.y);
    const V2 mp0 = mk(r1.x, r1.y), mp1 = mk(r2.y, r2.z);
    if (type == MANIFOLD_CIRCLES) {
      normal = mk(1.0f, 0.0f);
      V2 pointA = xmul(xfA, lp);
      V2 pointB = xmul(xfB, mp0);
      if (distsq(pointA, pointB) > KB_EPS * KB_EPS) {
        normal = pointB - pointA;
        normalize(normal);
      }
      V2 a = pointA + radiusA * normal;
      V2 b = pointB - radiusB * normal;
      pts[0] = 0.5f * (a + b);
      pts[1] = pts[0];
    } else if (type == MANIFOLD_FACE_A) {
      normal = rmul(xfA.q, ln);
      V2 planePoint = xmul(xfA, lp);
      for (int k = 0; k < pointCount; ++k) {
        V2 clipPoint = xmul(xfB, k == 0 ? mp0 : mp1);
        V2 a = clipPoint + (radiusA - dot(clipPoint - planePoint, normal)) * normal;
        V2 b = clipPoint - radiusB * normal;
        pts[k] = 0.5f * (a + b);
      }
    } else {
      normal = rmul(xfB.q, ln);
      V2 planePoint = xmul(xfB, lp);
      for (int k = 0; k < pointCount; ++k) {
        V2 clipPoint = xmul(xfA, k == 0 ? mp0 : mp1);
        V2 b = clipPoint + (radiusB - dot(clipPoint - planePoint, normal)) * normal;
        V2 a = clipPoint - radiusA * normal;
        pts[k] = 0.5f * (a + b);
      }
      normal = -normal;
    }
    const float friction = sqrtf(__ldg(&px[pa].friction) * __ldg(&px[pb].friction));
    const float restA = __ldg(&px[pa].restitution), restB = __ldg(&px[pb].restitution);
    const float restitution = restA > restB ? restA : restB;
    const V2 tangent = cross(normal, 1.0f);
    float nMass[2], tMass[2], bias[2];
    V2 rA[2], rB[2];
    for (int k = 0; k < pointCount; ++k) {
      rA[k] = pts[k] - cA;
      rB[k] = pts[k] - cB;
      float rnA = cross(rA[k], normal);
      float rnB = cross(rB[k], normal);
      float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
      nMass[k] = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
      float rtA = cross(rA[k], tangent);
      float rtB = cross(rB[k], tangent);
      float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
      tMass[k] = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
      bias[k] = 0.0f;
      float vRel = dot(normal, vB + cross(wB, rB[k]) - vA - cross(wA, rA[k]));
      if (vRel < -KB_VELOCITY_THRESHOLD) bias[k] = -restitution * vRel;
    }
    int vcPointCount = pointCount;
    float k11 = 0.0f, k12 = 0.0f, k22 = 0.0f, nm00 = 0.0f, nm01 = 0.0f, nm10 = 0.0f, nm11 = 0.0f;
    if (pointCount == 2) {
      float rn1A = cross(rA[0], normal);
      float rn1B = cross(rB[0], normal);
      float rn2A = cross(rA[1], normal);
      float rn2B = cross(rB[1], normal);
      k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
      k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
      k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
      const float k_maxConditionNumber = 1000.0f;
      if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
        // b2Mat22::GetInverse with ex = (k11, k12), ey = (k12, k22)
        float a = k11, b = k12, c = k12, d = k22;
        float det = a * d - b * c;
        if (det != 0.0f) det = 1.0f / det;
        nm00 = det * d;    // ex.x
        nm01 = -det * b;   // ey.x
        nm10 = -det * c;   // ex.y
        nm11 = det * a;    // ey.y
      } else {
        vcPointCount = 1;
      }
    }
    poolu(PF_IDX, q) = (uint32_t)bA | ((uint32_t)bB << 8) | ((uint32_t)vcPointCount << 16) |
                       ((uint32_t)(general ? 1 : 0) << 20) | ((uint32_t)type << 24) | ((uint32_t)pointCount << 28);
    poolu(PF_AUX, q) = oc;
    pool(PF_NX, q) = normal.x;
    pool(PF_NY, q) = normal.y;
    pool(PF_RAX, q) = rA[0].x;
    pool(PF_RAY, q) = rA[0].y;
    pool(PF_RBX, q) = rB[0].x;
    pool(PF_RBY, q) = rB[0].y;
    pool(PF_NMASS, q) = nMass[0];
    pool(PF_NIMP, q) = fresh ? 0.0f : r1.z;  // warm start: dtRatio (== 1) * normalImpulse
    if (general) {
      gw(q, 10) = tMass[0];
      gw(q, 11) = fresh ? 0.0f : r1.w;  // tangentImpulse
      gw(q, 12) = bias[0];
      gw(q, 13) = friction;
      gw(q, 14) = k11;
      gw(q, 15) = k12;
      gw(q, 16) = k22;
      gw(q, 17) = nm00;
      gw(q, 18) = nm01;
      gw(q, 19) = nm11;
      if (pointCount == 2) {
        gw(q, 20) = rA[1].x;
        gw(q, 21) = rA[1].y;
        gw(q, 22) = rB[1].x;
        gw(q, 23) = rB[1].y;
        gw(q, 24) = nMass[1];
        gw(q, 25) = fresh ? 0.0f : r2.w;  // p1 normalImpulse
        gw(q, 26) = tMass[1];
        gw(q, 27) = fresh ? 0.0f : r3.x;  // p1 tangentImpulse
        gw(q, 28) = bias[1];
      }
      gw(q, 29) = nm10;
    }
  }

  __device__ __forceinline__ void loadVel(int b, VelBody& o) {
    float4 v = vel4(b);
    o.v = mk(v.x, v.y);
    o.w = v.z;
  }
  __device__ __forceinline__ void storeVel(int b, const VelBody& o) {
    if (b == S) return;
    float* v = reinterpret_cast<float*>(&vel4(b));
    v[0] = o.v.x;
    v[1] = o.v.y;
    v[2] = o.w;
  }

  // b2ContactSolver::WarmStart for one constraint
  __device__ __forceinline__ void warmStartOne(int q) {
    const uint32_t idx = poolu(PF_IDX, q);
    const int bA = idx & 0xFF, bB = (idx >> 8) & 0xFF;
    const int pointCount = (idx >> 16) & 0xF;
    const bool general = ((idx >> 20) & 1u) != 0u;
    const float4 kA = bc4(bA), kB = bc4(bB);
    const float mA = kA.x, iA = kA.y, mB = kB.x, iB = kB.y;
    VelBody A, Bv;
    loadVel(bA, A);
    loadVel(bB, Bv);
    const V2 normal = mk(pool(PF_NX, q), pool(PF_NY, q));
    const V2 tangent = cross(normal, 1.0f);
    {
      const V2 rA = mk(pool(PF_RAX, q), pool(PF_RAY, q)), rB = mk(pool(PF_RBX, q), pool(PF_RBY, q));
      const float ni = pool(PF_NIMP, q);
      const float ti = general ? gw(q, 11) : 0.0f;
      V2 P = ni * normal + ti * tangent;
      A.w -= iA * cross(rA, P);
      A.v = A.v - mA * P;
      Bv.w += iB * cross(rB, P);
      Bv.v = Bv.v + mB * P;
    }
    if (pointCount == 2) {
      const V2 rA = mk(gw(q, 20), gw(q, 21)), rB = mk(gw(q, 22), gw(q, 23));
      V2 P = gw(q, 25) * normal + gw(q, 27) * tangent;
      A.w -= iA * cross(rA, P);
      A.v = A.v - mA * P;
      Bv.w += iB * cross(rB, P);
      Bv.v = Bv.v + mB * P;
    }
    storeVel(bA, A);
    storeVel(bB, Bv);
  }

  // b2ContactSolver::SolveVelocityConstraints for one constraint
  __device__ __forceinline__ void solveVelocityOne(int q) {
    const uint32_t idx = poolu(PF_IDX, q);
    const int bA = idx & 0xFF, bB = (idx >> 8) & 0xFF;
    const int pointCount = (idx >> 16) & 0xF;
    const bool general = ((idx >> 20) & 1u) != 0u;
    const float4 kA = bc4(bA), kB = bc4(bB);
    const float mA = kA.x, iA = kA.y, mB = kB.x, iB = kB.y;
    VelBody A, Bv;
    loadVel(bA, A);
    loadVel(bB, Bv);
    const V2 normal = mk(pool(PF_NX, q), pool(PF_NY, q));
    const V2 rA0 = mk(pool(PF_RAX, q), pool(PF_RAY, q)), rB0 = mk(pool(PF_RBX, q), pool(PF_RBY, q));
    if (!general) {
      // frictionless single point: the tangent row contributes exactly zero (friction == 0)
      V2 dv = Bv.v + cross(Bv.w, rB0) - A.v - cross(A.w, rA0);
      float vn = dot(dv, normal);
      float ni = pool(PF_NIMP, q);
      float lambda = -pool(PF_NMASS, q) * (vn - 0.0f);
      float newImpulse = b2max(ni + lambda, 0.0f);
      lambda = newImpulse - ni;
      pool(PF_NIMP, q) = newImpulse;
      V2 P = lambda * normal;
      A.v = A.v - mA * P;
      A.w -= iA * cross(rA0, P);
      Bv.v = Bv.v + mB * P;
      Bv.w += iB * cross(rB0, P);
    } else {
      const V2 tangent = cross(normal, 1.0f);
      const float friction = gw(q, 13);
      V2 rA1 = rA0, rB1 = rB0;
      if (pointCount == 2) {
        rA1 = mk(gw(q, 20), gw(q, 21));
        rB1 = mk(gw(q, 22), gw(q, 23));
      }
      // tangent rows first
      for (int j = 0; j < pointCount; ++j) {
        const V2 rA = j == 0 ? rA0 : rA1, rB = j == 0 ? rB0 : rB1;
        const float tMass = j == 0 ? gw(q, 10) : gw(q, 26);
        const float nImp = j == 0 ? pool(PF_NIMP, q) : gw(q, 25);
        float& tImp = j == 0 ? gw(q, 11) : gw(q, 27);
        V2 dv = Bv.v + cross(Bv.w, rB) - A.v - cross(A.w, rA);
        float vt = dot(dv, tangent) - 0.0f;
        float lambda = tMass * (-vt);
        float maxFriction = friction * nImp;
        float newImpulse = b2clamp(tImp + lambda, -maxFriction, maxFriction);
        lambda = newImpulse - tImp;
        tImp = newImpulse;
        V2 P = lambda * tangent;
        A.v = A.v - mA * P;
        A.w -= iA * cross(rA, P);
        Bv.v = Bv.v + mB * P;
        Bv.w += iB * cross(rB, P);
      }
      if (pointCount == 1) {
        V2 dv = Bv.v + cross(Bv.w, rB0) - A.v - cross(A.w, rA0);
        float vn = dot(dv, normal);
        float ni = pool(PF_NIMP, q);
        float lambda = -pool(PF_NMASS, q) * (vn - gw(q, 12));
        float newImpulse = b2max(ni + lambda, 0.0f);
        lambda = newImpulse - ni;
        pool(PF_NIMP, q) = newImpulse;
        V2 P = lambda * normal;
        A.v = A.v - mA * P;
        A.w -= iA * cross(rA0, P);
        Bv.v = Bv.v + mB * P;
        Bv.w += iB * cross(rB0, P);
      } else {
        // block solver
        const float k11 = gw(q, 14), k12 = gw(q, 15), k22 = gw(q, 16);
        const float nm00 = gw(q, 17), nm01 = gw(q, 18), nm10 = gw(q, 29), nm11 = gw(q, 19);
        const float nMass1 = pool(PF_NMASS, q), nMass2 = gw(q, 24);
        V2 a = mk(pool(PF_NIMP, q), gw(q, 25));
        V2 dv1 = Bv.v + cross(Bv.w, rB0) - A.v - cross(A.w, rA0);
        V2 dv2 = Bv.v + cross(Bv.w, rB1) - A.v - cross(A.w, rA1);
        float vn1 = dot(dv1, normal);
        float vn2 = dot(dv2, normal);
        V2 b;
        b.x = vn1 - gw(q, 12);
        b.y = vn2 - gw(q, 28);
        // b -= K a ; K.ex = (k11,k12), K.ey = (k12,k22)
        b = b - mk(k11 * a.x + k12 * a.y, k12 * a.x + k22 * a.y);
        V2 x;
        bool found = false;
        // case 1
        {
          V2 t = mk(nm00 * b.x + nm01 * b.y, nm10 * b.x + nm11 * b.y);
          x = -t;
          if (x.x >= 0.0f && x.y >= 0.0f) found = true;
        }
        if (!found) {  // case 2
          x.x = -nMass1 * b.x;
          x.y = 0.0f;
          vn2 = k12 * x.x + b.y;
          if (x.x >= 0.0f && vn2 >= 0.0f) found = true;
        }
        if (!found) {  // case 3
          x.x = 0.0f;
          x.y = -nMass2 * b.y;
          vn1 = k12 * x.y + b.x;
          if (x.y >= 0.0f && vn1 >= 0.0f) found = true;
        }
        if (!found) {  // case 4
          x.x = 0.0f;
          x.y = 0.0f;
          vn1 = b.x;
          vn2 = b.y;
          if (vn1 >= 0.0f && vn2 >= 0.0f) found = true;
        }
        if (found) {
          V2 d = x - a;
          V2 P1 = d.x * normal;
          V2 P2 = d.y * normal;
          A.v = A.v - mA * (P1 + P2);
          A.w -= iA * (cross(rA0, P1) + cross(rA1, P2));
          Bv.v = Bv.v + mB * (P1 + P2);
          Bv.w += iB * (cross(rB0, P1) + cross(rB1, P2));
          pool(PF_NIMP, q) = x.x;
          gw(q, 25) = x.y;
        }
      }
    }
    storeVel(bA, A);
    storeVel(bB, Bv);
  }

  // b2ContactSolver::StoreImpulses, then re-purpose the pool for the position solver
  __device__ __forceinline__ void storeImpulsesAndPreparePosition(int e) {
    const uint32_t en = entry(e);
    const int q = en >> 16;
    const uint32_t idx = poolu(PF_IDX, q);
    const uint32_t aux = poolu(PF_AUX, q);
    const int ci = aux & 0xFFFF;
    const int vcPointCount = (idx >> 16) & 0xF;
    const bool general = ((idx >> 20) & 1u) != 0u;
    float* rec = manifoldRec(ci);
    rec[MR_P0N] = pool(PF_NIMP, q);
    if (general) {
      rec[MR_P0T] = gw(q, 11);
      if (vcPointCount == 2) {
        rec[MR_P1N] = gw(q, 25);
        rec[MR_P1T] = gw(q, 27);
      }
    }
    // position data: localNormal, localPoint, radii (+ both local points for general contacts)
    const float4 r0 = reinterpret_cast<const float4*>(rec)[0];
    const uint32_t pr = cpair(ci);
    const int pa = pr & 0xFFFF, pb = pr >> 16;
    pool(PF_NX, q) = r0.x;
    pool(PF_NY, q) = r0.y;
    pool(PF_RAX, q) = r0.z;
    pool(PF_RAY, q) = r0.w;
    pool(PF_RBX, q) = __ldg(&px[pa].radius);
    pool(PF_RBY, q) = __ldg(&px[pb].radius);
    if (general) {
      const float4 r1 = reinterpret_cast<const float4*>(rec)[1];
      const float4 r2 = reinterpret_cast<const float4*>(rec)[2];
      gw(q, 10) = r1.x;
      gw(q, 11) = r1.y;
      gw(q, 12) = r2.y;
      gw(q, 13) = r2.z;
    }
  }

  // one constraint of b2ContactSolver::SolvePositionConstraints / SolveTOIPositionConstraints.
  // Returns false if any point's separation is below `limit` (island not yet solved).
  __device__ __forceinline__ bool solvePositionOne(int q, float baumgarte, float limit, int toiA, int toiB) {
    const uint32_t idx = poolu(PF_IDX, q);
    const int bA = idx & 0xFF, bB = (idx >> 8) & 0xFF;
    const bool general = ((idx >> 20) & 1u) != 0u;
    const int type = (idx >> 24) & 0xF;
    const int pointCount = (idx >> 28) & 0xF;
    const float4 kA = bc4(bA), kB = bc4(bB);
    float mA = kA.x, iA = kA.y, mB = kB.x, iB = kB.y;
    if (toiA >= 0) {
      if (!(bA == toiA || bA == toiB)) { mA = 0.0f; iA = 0.0f; }
      if (!(bB == toiA || bB == toiB)) { mB = 0.0f; iB = 0.0f; }
    }
    const V2 lcA = mk(kA.z, kA.w), lcB = mk(kB.z, kB.w);
    float4 pA4 = pos4(bA), pB4 = pos4(bB);
    V2 cA = mk(pA4.x, pA4.y), cB = mk(pB4.x, pB4.y);
    float aA = pA4.z, aB = pB4.z;
    const V2 ln = mk(pool(PF_NX, q), pool(PF_NY, q));
    const V2 lp = mk(pool(PF_RAX, q), pool(PF_RAY, q));
    const float radiusA = pool(PF_RBX, q), radiusB = pool(PF_RBY, q);
    // a rotation is only needed where it multiplies something non-zero (0 * finite == 0 exactly)
    const bool trigA = bA != S && (type != MANIFOLD_CIRCLES || lp.x != 0.0f || lp.y != 0.0f || lcA.x != 0.0f ||
                                   lcA.y != 0.0f || general);
    const bool trigB = bB != S && (type == MANIFOLD_FACE_B || lcB.x != 0.0f || lcB.y != 0.0f || general);
    bool ok = true;
    for (int j = 0; j < pointCount; ++j) {
      Xf xfA, xfB;
      if (trigA) xfA.q = rot_set(aA);
      else { xfA.q.s = 0.0f; xfA.q.c = 1.0f; }
      if (trigB) xfB.q = rot_set(aB);
      else { xfB.q.s = 0.0f; xfB.q.c = 1.0f; }
      xfA.p = cA - rmul(xfA.q, lcA);
      xfB.p = cB - rmul(xfB.q, lcB);
      V2 mpj = mk(0.0f, 0.0f);
      if (general) mpj = j == 0 ? mk(gw(q, 10), gw(q, 11)) : mk(gw(q, 12), gw(q, 13));
      V2 normal, point;
      float separation;
      if (type == MANIFOLD_CIRCLES) {
        V2 pointA = xmul(xfA, lp);
        V2 pointB = xmul(xfB, mpj);
        normal = pointB - pointA;
        normalize(normal);
        point = 0.5f * (pointA + pointB);
        separation = dot(pointB - pointA, normal) - radiusA - radiusB;
      } else if (type == MANIFOLD_FACE_A) {
        normal = rmul(xfA.q, ln);
        V2 planePoint = xmul(xfA, lp);
        V2 clipPoint = xmul(xfB, mpj);
        separation = dot(clipPoint - planePoint, normal) - radiusA - radiusB;
        point = clipPoint;
      } else {
        normal = rmul(xfB.q, ln);
        V2 planePoint = xmul(xfB, lp);
        V2 clipPoint = xmul(xfA, mpj);
        separation = dot(clipPoint - planePoint, normal) - radiusA - radiusB;
        point = clipPoint;
        normal = -normal;
      }
      V2 rA = point - cA;
      V2 rB = point - cB;
      if (!(separation >= limit)) ok = false;
      float C = b2clamp(baumgarte * (separation + KB_LINEAR_SLOP), -KB_MAX_LINEAR_CORRECTION, 0.0f);
      float rnA = cross(rA, normal);
      float rnB = cross(rB, normal);
      float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
      float impulse = K > 0.0f ? -C / K : 0.0f;
      V2 P = impulse * normal;
      cA = cA - mA * P;
      aA -= iA * cross(rA, P);
      cB = cB + mB * P;
      aB += iB * cross(rB, P);
    }
    if (bA != S) {
      float* o = reinterpret_cast<float*>(&pos4(bA));
      o[0] = cA.x; o[1] = cA.y; o[2] = aA;
    }
    if (bB != S) {
      float* o = reinterpret_cast<float*>(&pos4(bB));
      o[0] = cB.x; o[1] = cB.y; o[2] = aB;
    }
    return ok;
  }

  // b2World::Solve
  __device__ void solve() {
    const int nC = (int)hdr(H_NC);
    const int B = L.B;
    unsigned long long* cnt = counters();
    // ---- touching list in world-list order (descending index)
    uint32_t* tlist = sm + L.sTlist;
    int K = 0;
    for (int base = 0; base < nC; base += LPE) {
      const int i = nC - 1 - (base + g.lane);
      bool t = false;
      uint32_t val = 0u;
      if (i >= 0) {
        const uint32_t info = cinfo(i);
        t = (info & (CI_TOUCHING | CI_ENABLED)) == (CI_TOUCHING | CI_ENABLED);
        if (t) {
          const uint32_t pr = cpair(i);
          const int bA = __ldg(&px[pr & 0xFFFF].body), bB = __ldg(&px[pr >> 16].body);
          val = (uint32_t)i | ((uint32_t)bA << 16) | ((uint32_t)bB << 24);
        }
      }
      const uint32_t m = g.ballot(t);
      const int dst = K + __popc(m & g.lt());
      if (t && dst < L.Kmax) tlist[dst] = val;
      K += __popc(m);
    }
    if (K > L.Kmax) {
      if (g.lane == 0) hdr(H_STATUS) |= KB_STATUS_SOLVER_OVERFLOW;
      K = L.Kmax;
    }
    for (int b = g.lane; b <= B; b += LPE) isl(b) = -1;
    g.sync();
    // ---- island DFS (b2World::Solve).  bodies flagged in (flo, fhi); contacts flagged per lane.
    uint32_t flo = 0u, fhi = 0u;  // island flags of bodies, replicated in every lane
    uint32_t cflag = 0u;          // bit s: tlist entry (s * LPE + lane) already in an island
    int nOrd = 0, nIslands = 0;
    int32_t* stack = reinterpret_cast<int32_t*>(sm + L.sStack);
    const int chunks = (K + LPE - 1) / LPE;
    for (int seed = B - 1; seed >= 0; --seed) {
      if (((seed < 32 ? flo >> seed : fhi >> (seed - 32)) & 1u) != 0u) continue;
      if (!awake(seed)) continue;
      int sp = 0;
      if (g.lane == 0) stack[0] = seed;
      sp = 1;
      if (seed < 32) flo |= 1u << seed; else fhi |= 1u << (seed - 32);
      g.sync();
      while (sp > 0) {
        const int b = stack[sp - 1];
        --sp;
        g.sync();
        if (g.lane == 0) {
          isl(b) = nIslands;
          wake(b);
        }
        for (int s = 0; s < chunks; ++s) {
          const int t = s * LPE + g.lane;
          bool inv = false;
          uint32_t tv = 0u;
          int other = S;
          if (t < K && ((cflag >> s) & 1u) == 0u) {
            tv = tlist[t];
            const int bA = (tv >> 16) & 0xFF, bB = tv >> 24;
            inv = bA == b || bB == b;
            other = bA == b ? bB : bA;
          }
          const uint32_t m = g.ballot(inv);
          if (m == 0u) continue;
          if (inv) {
            const int dst = nOrd + __popc(m & g.lt());
            ordC(dst) = (tv & 0xFFFF) | ((uint32_t)nIslands << 16);
            cflag |= 1u << s;
          }
          nOrd += __popc(m);
          // push unflagged dynamic neighbours in list order, first occurrence only
          const bool cand = inv && other != S && ((other < 32 ? flo >> other : fhi >> (other - 32)) & 1u) == 0u;
          const uint32_t same = g.match(cand ? (uint32_t)other : (0x100u + (uint32_t)g.lane));
          const bool first = cand && (__ffs(same) - 1) == g.lane;
          const uint32_t pm = g.ballot(first);
          if (first) stack[sp + __popc(pm & g.lt())] = other;
          sp += __popc(pm);
          const uint32_t addlo = g.red_or(first && other < 32 ? 1u << other : 0u);
          const uint32_t addhi = g.red_or(first && other >= 32 ? 1u << (other - 32) : 0u);
          flo |= addlo;
          fhi |= addhi;
          g.sync();
        }
      }
      ++nIslands;
    }
    g.sync();
    if (g.lane == 0) {
      cnt[KB_CNT_ISLANDS] += (unsigned long long)nIslands;
    }
    // ---- classify ordered contacts (simple: 1 slot, general: 3 slots) and fetch body ids
    for (int p = g.lane; p < nOrd; p += LPE) {
      const int ci = ordC(p) & 0xFFFF;
      const uint32_t pr = cpair(ci);
      const int pa = pr & 0xFFFF, pb = pr >> 16;
      const int bA = __ldg(&px[pa].body), bB = __ldg(&px[pb].body);
      const int pc = (cinfo(ci) & CI_PC_MASK) >> CI_PC_SHIFT;
      const float fr = __ldg(&px[pa].friction) * __ldg(&px[pb].friction);
      const float re = b2max(__ldg(&px[pa].restitution), __ldg(&px[pb].restitution));
      const bool simple = pc == 1 && fr == 0.0f && re == 0.0f && __ldg(&px[pb].type) == SHAPE_CIRCLE;
      ordB(p) = (uint32_t)bA | ((uint32_t)bB << 8) | ((simple ? 1u : 3u) << 16) | ((uint32_t)pc << 24);
    }
    g.sync();
    // ---- dependency levels + counting sort by level (serial, lane 0)
    int numLevels = 0, nEntries = 0;
    if (g.lane == 0) {
      uint32_t* lastLvl = sm + L.sLastLvl;
      uint32_t* lvl = sm + L.sLvl;  // level per order position
      for (int b = 0; b <= B; ++b) lastLvl[b] = 0u;
      int maxL = 0;
      for (int p = 0; p < nOrd; ++p) {
        const uint32_t ob = ordB(p);
        const int bA = ob & 0xFF, bB = (ob >> 8) & 0xFF;
        uint32_t l = 0u;
        if (bA != S) l = lastLvl[bA];
        if (bB != S) l = max(l, lastLvl[bB]);
        l += 1u;
        if (bA != S) lastLvl[bA] = l;
        if (bB != S) lastLvl[bB] = l;
        lvl[p] = l;
        maxL = max(maxL, (int)l);
      }
      // count entries / slots per level (lvlOff doubles as histogram)
      uint32_t* slotOff = sm + L.sLvlOff + L.Kmax + 2;
      for (int l = 0; l <= maxL + 1; ++l) {
        lvlOff(l) = 0u;
        slotOff[l] = 0u;
      }
      for (int p = 0; p < nOrd; ++p) {
        lvlOff(lvl[p] + 1) += 1u;
        slotOff[lvl[p] + 1] += (ordB(p) >> 16) & 0xFF;
      }
      for (int l = 1; l <= maxL + 1; ++l) {
        lvlOff(l) += lvlOff(l - 1);
        slotOff[l] += slotOff[l - 1];
      }
      // lvlOff(l) = first entry of level l (levels are 1-based; lvlOff(maxL+1) = nOrd)
      int dropped = 0;
      for (int p = 0; p < nOrd; ++p) {
        const uint32_t l = lvl[p];
        const uint32_t size = (ordB(p) >> 16) & 0xFF;
        // cursors: lvlOff(l) and slotOff[l] are advanced in place and restored afterwards
        const uint32_t e = lvlOff(l);
        const uint32_t q = slotOff[l];
        lvlOff(l) = e + 1u;
        slotOff[l] = q + size;
        if ((int)(q + size) > L.Kmax) {
          entry(e) = 0xFFFFFFFFu;
          ++dropped;
        } else {
          entry(e) = (uint32_t)p | (q << 16);
        }
      }
      // restore offsets: after the pass lvlOff(l) == start of level l+1
      for (int l = maxL + 1; l >= 1; --l) lvlOff(l) = lvlOff(l - 1);
      lvlOff(0) = 0u;
      if (dropped) hdr(H_STATUS) |= KB_STATUS_SOLVER_OVERFLOW;
      misc(0) = (uint32_t)maxL;
      cnt[KB_CNT_LEVELS] += (unsigned long long)maxL;
    }
    g.sync();
    numLevels = (int)misc(0);
    nEntries = nOrd;
    // ---- b2Island::Solve: integrate velocities (damping), remember the sweep start
    const float h = L.dt;
    for (int b = g.lane; b < B; b += LPE) {
      if (isl(b) < 0) continue;
      const float4 p = pos4(b);
      const float4 x = xf4(b);
      sweep4(b) = make_float4(p.x, p.y, p.z, 0.0f);
      reinterpret_cast<float2*>(sm + L.sSweep + 4 * (L.B + 1))[b] = make_float2(x.z, x.w);
      float4 v = vel4(b);
      // (SimplePhototaxisKilobot's linearDamping = 0, lib/kilobot.py:203, is folded into the template)
      const float ld = __ldg(&bc[b].linearDamping);
      const float ad = __ldg(&bc[b].angularDamping);
      if (L.dampingMode == 0) {
        const float fl = 1.0f / (1.0f + h * ld);
        const float fa = 1.0f / (1.0f + h * ad);
        v.x *= fl;
        v.y *= fl;
        v.z *= fa;
      } else {
        const float fl = b2clamp(1.0f - h * ld, 0.0f, 1.0f);
        const float fa = b2clamp(1.0f - h * ad, 0.0f, 1.0f);
        v.x *= fl;
        v.y *= fl;
        v.z *= fa;
      }
      vel4(b) = v;
    }
    g.sync();
    // ---- constraints
    for (int e = g.lane; e < nEntries; e += LPE)
      if (entry(e) != 0xFFFFFFFFu) initConstraint(e);
    g.sync();
    unsigned long long pts = 0ull;
    for (int e = g.lane; e < nEntries; e += LPE)
      if (entry(e) != 0xFFFFFFFFu) pts += (ordB(entry(e) & 0xFFFF) >> 24) & 0xF;
    pts = (unsigned long long)g.red_add((uint32_t)pts);
    if (g.lane == 0) cnt[KB_CNT_POINTS] += pts;
    // warm start (ordered)
    for (int l = 1; l <= numLevels; ++l) {
      const int e0 = (int)lvlOff(l), e1 = (int)lvlOff(l + 1);
      for (int eb = e0; eb < e1; eb += LPE) {
        const int e = eb + g.lane;
        if (e < e1 && entry(e) != 0xFFFFFFFFu) warmStartOne(entry(e) >> 16);
      }
      g.sync();
    }
    for (int it = 0; it < L.velIters; ++it) {
      for (int l = 1; l <= numLevels; ++l) {
        const int e0 = (int)lvlOff(l), e1 = (int)lvlOff(l + 1);
        for (int eb = e0; eb < e1; eb += LPE) {
          const int e = eb + g.lane;
          if (e < e1 && entry(e) != 0xFFFFFFFFu) solveVelocityOne(entry(e) >> 16);
        }
        g.sync();
      }
    }
    for (int e = g.lane; e < nEntries; e += LPE)
      if (entry(e) != 0xFFFFFFFFu) storeImpulsesAndPreparePosition(e);
    // ---- integrate positions
    for (int b = g.lane; b < B; b += LPE) {
      if (isl(b) < 0) continue;
      float4 p = pos4(b);
      float4 v = vel4(b);
      V2 translation = h * mk(v.x, v.y);
      if (dot(translation, translation) > KB_MAX_TRANSLATION_SQ) {
        float ratio = KB_MAX_TRANSLATION / length(translation);
        v.x *= ratio;
        v.y *= ratio;
      }
      float rotation = h * v.z;
      if (rotation * rotation > KB_MAX_ROTATION_SQ) {
        float ratio = KB_MAX_ROTATION / b2abs(rotation);
        v.z *= ratio;
      }
      p.x += h * v.x;
      p.y += h * v.y;
      p.z += h * v.z;
      pos4(b) = p;
      vel4(b) = v;
    }
    for (int i = g.lane; i < nIslands; i += LPE) islflag(i) = 0u;  // bit0: unsolved this iteration, bit1: solved
    g.sync();
    // ---- position iterations with per-island early exit
    {
      int remaining = nIslands;
      for (int it = 0; it < L.posIters && remaining > 0; ++it) {
        if (g.lane == 0) cnt[KB_CNT_POS_ITERS] += (unsigned long long)remaining;
        for (int l = 1; l <= numLevels; ++l) {
          const int e0 = (int)lvlOff(l), e1 = (int)lvlOff(l + 1);
          for (int eb = e0; eb < e1; eb += LPE) {
            const int e = eb + g.lane;
            if (e < e1 && entry(e) != 0xFFFFFFFFu) {
              const int q = entry(e) >> 16;
              const int island = poolu(PF_AUX, q) >> 16;
              if ((islflag(island) & 2u) == 0u) {
                const bool ok = solvePositionOne(q, KB_BAUMGARTE, -3.0f * KB_LINEAR_SLOP, -1, -1);
                if (!ok) atomicOr(&islflag(island), 1u);
              }
            }
          }
          g.sync();
        }
        int solvedNow = 0;
        for (int i = g.lane; i < nIslands; i += LPE) {
          const uint32_t f = islflag(i);
          if ((f & 2u) == 0u) {
            if ((f & 1u) == 0u) {
              islflag(i) = 2u;
              ++solvedNow;
            } else {
              islflag(i) = 0u;
            }
          }
        }
        remaining -= (int)g.red_add((uint32_t)solvedNow);
        g.sync();
      }
    }
    // ---- copy back: SynchronizeTransform; sleep bookkeeping
    for (int b = g.lane; b < B; b += LPE) {
      const int island = isl(b);
      if (island < 0) continue;
      const float4 p = pos4(b);
      const float4 k = bc4(b);
      Rot q = rot_set(p.z);
      V2 o = mk(p.x, p.y) - rmul(q, mk(k.z, k.w));
      xf4(b) = make_float4(o.x, o.y, q.s, q.c);
      if (L.enableSleep) {
        const float4 v = vel4(b);
        const float linTolSqr = KB_LIN_SLEEP_TOL * KB_LIN_SLEEP_TOL;
        const float angTolSqr = KB_ANG_SLEEP_TOL * KB_ANG_SLEEP_TOL;
        float st;
        if (v.z * v.z > angTolSqr || dot(mk(v.x, v.y), mk(v.x, v.y)) > linTolSqr) {
          st = 0.0f;
        } else {
          st = p.w + h;
        }
        reinterpret_cast<float*>(&pos4(b))[3] = st;
        if (!(st >= KB_TIME_TO_SLEEP)) atomicOr(&islflag(island), 4u);  // minSleepTime < timeToSleep
      }
    }
    g.sync();
    if (L.enableSleep) {
      for (int b = g.lane; b < B; b += LPE) {
        const int island = isl(b);
        if (island < 0) continue;
        const uint32_t f = islflag(island);
        if ((f & 4u) == 0u && (f & 2u) != 0u) {
          // b2Body::SetAwake(false)
          vel4(b) = make_float4(0.0f, 0.0f, 0.0f, u2f(f2u(vel4(b).w) & ~BF_AWAKE));
          reinterpret_cast<float*>(&pos4(b))[3] = 0.0f;
        }
      }
      g.sync();
    }
    synchronizeFixtures(false);
    findNewContacts();
  }

  // shape AABB of proxy p under transform xf (b2Shape::ComputeAABB)
  __device__ __forceinline__ void shapeAABB(int p, Xf xf, V2* lo, V2* hi) {
    const ProxyConst* pc = px + p;
    const int type = __ldg(&pc->type);
    if (type == SHAPE_CIRCLE) {
      const float r = __ldg(&pc->radius);
      V2 c = xf.p + rmul(xf.q, mk(0.0f, 0.0f));
      *lo = mk(c.x - r, c.y - r);
      *hi = mk(c.x + r, c.y + r);
    } else if (type == SHAPE_POLYGON) {
      const int n = __ldg(&pc->count);
      V2 lower = xmul(xf, pvert(pc, 0));
      V2 upper = lower;
      for (int i = 1; i < n; ++i) {
        V2 v = xmul(xf, pvert(pc, i));
        lower = mk(b2min(lower.x, v.x), b2min(lower.y, v.y));
        upper = mk(b2max(upper.x, v.x), b2max(upper.y, v.y));
      }
      const float r = __ldg(&pc->radius);
      *lo = mk(lower.x - r, lower.y - r);
      *hi = mk(upper.x + r, upper.y + r);
    } else {
      V2 a = xmul(xf, pvert(pc, 1));
      V2 b = xmul(xf, pvert(pc, 2));
      *lo = mk(b2min(a.x, b.x), b2min(a.y, b.y));
      *hi = mk(b2max(a.x, b.x), b2max(a.y, b.y));
    }
  }

  // b2Body::SynchronizeFixtures + b2BroadPhase::MoveProxy for every body that was in an island
  // (toiMode: for bodies flagged in isl() by the TOI mini-island; xf1 is rebuilt from (c0, a0)).
  __device__ void synchronizeFixtures(bool toiMode) {
    uint32_t mlo = 0u, mhi = 0u;
    for (int p = g.lane; p < L.P; p += LPE) {
      const int b = __ldg(&px[p].body);
      if (b == S || b < 0) continue;
      if (isl(b) < 0) continue;
      const float4 sw = sweep4(b);
      const float4 k = bc4(b);
      Xf xf1;
      if (toiMode) {
        xf1.q = rot_set(sw.z);
      } else {
        const float2 oq = reinterpret_cast<const float2*>(sm + L.sSweep + 4 * (L.B + 1))[b];
        xf1.q.s = oq.x;
        xf1.q.c = oq.y;
      }
      xf1.p = mk(sw.x, sw.y) - rmul(xf1.q, mk(k.z, k.w));
      const Xf xf2 = bodyXf(b);
      V2 lo1, hi1, lo2, hi2;
      shapeAABB(p, xf1, &lo1, &hi1);
      shapeAABB(p, xf2, &lo2, &hi2);
      const V2 lo = mk(b2min(lo1.x, lo2.x), b2min(lo1.y, lo2.y));
      const V2 hi = mk(b2max(hi1.x, hi2.x), b2max(hi1.y, hi2.y));
      const V2 displacement = xf2.p - xf1.p;
      const float4 fat = fat4(p);
      const bool contains = fat.x <= lo.x && fat.y <= lo.y && hi.x <= fat.z && hi.y <= fat.w;
      if (!contains) {
        float4 nb = make_float4(lo.x - KB_AABB_EXTENSION, lo.y - KB_AABB_EXTENSION, hi.x + KB_AABB_EXTENSION,
                                hi.y + KB_AABB_EXTENSION);
        const V2 d = KB_AABB_MULTIPLIER * displacement;
        if (d.x < 0.0f) nb.x += d.x; else nb.z += d.x;
        if (d.y < 0.0f) nb.y += d.y; else nb.w += d.y;
        fat4(p) = nb;
        if (p < 32) mlo |= 1u << p; else mhi |= 1u << (p - 32);
      }
    }
    mlo = g.red_or(mlo);
    mhi = g.red_or(mhi);
    if (g.lane == 0) {
      sm[L.sMoved] |= mlo;
      sm[L.sMoved + 1] |= mhi;
    }
    g.sync();
  }

  // b2BroadPhase::UpdatePairs + b2ContactManager::AddPair.  Pairs (i < j) with a moved member are
  // visited in (i, j) order == Box2D's sorted pair buffer, so creation order matches.
  __device__ void findNewContacts() {
    const uint32_t mlo = sm[L.sMoved], mhi = sm[L.sMoved + 1];
    g.sync();
    if (g.lane == 0) {
      sm[L.sMoved] = 0u;
      sm[L.sMoved + 1] = 0u;
    }
    if ((mlo | mhi) == 0u) {
      g.sync();
      return;
    }
    const unsigned long long moved = (unsigned long long)mlo | ((unsigned long long)mhi << 32);
    int nC = (int)hdr(H_NC);
    bool overflow = false;
    unsigned long long tests = 0ull;
    const int P = L.P;
    for (int i = 0; i < P - 1; ++i) {
      const bool movedI = ((moved >> i) & 1ull) != 0ull;
      if (!movedI && (moved >> (i + 1)) == 0ull) break;
      const float4 fi = fat4(i);
      const int bi = __ldg(&px[i].body);
      const int ti = __ldg(&px[i].type);
      const unsigned long long adjI = (unsigned long long)adj(i, 0) | ((unsigned long long)adj(i, 1) << 32);
      for (int jb = i + 1; jb < P; jb += LPE) {
        const int j = jb + g.lane;
        bool create = false;
        int bj = S, tj = 0;
        if (j < P && (movedI || ((moved >> j) & 1ull) != 0ull)) {
          ++tests;
          const float4 fj = fat4(j);
          const bool overlap = !(fj.x - fi.z > 0.0f || fj.y - fi.w > 0.0f || fi.x - fj.z > 0.0f || fi.y - fj.w > 0.0f);
          bj = __ldg(&px[j].body);
          tj = __ldg(&px[j].type);
          create = overlap && bi != bj && ((adjI >> j) & 1ull) == 0ull;
        }
        const uint32_t m = g.ballot(create);
        if (m == 0u) continue;
        const int dst = nC + __popc(m & g.lt());
        if (create) {
          if (dst < L.Cmax) {
            // type register: chain edge < polygon < circle takes the A slot (b2Contact::Create)
            const int rankI = ti == SHAPE_EDGE ? 0 : (ti == SHAPE_POLYGON ? 1 : 2);
            const int rankJ = tj == SHAPE_EDGE ? 0 : (tj == SHAPE_POLYGON ? 1 : 2);
            const int pa = rankI > rankJ ? j : i, pb = rankI > rankJ ? i : j;
            cpair(dst) = (uint32_t)pa | ((uint32_t)pb << 16);
            cinfo(dst) = CI_ENABLED;
            atomicOr(&adj(i, j >> 5), 1u << (j & 31));
            atomicOr(&adj(j, i >> 5), 1u << (i & 31));
            wake(bj);
          } else {
            overflow = true;
          }
        }
        if (g.lane == 0 && bi != S) wake(bi);
        nC = min(nC + __popc(m), L.Cmax);
        g.sync();
      }
    }
    tests = (unsigned long long)g.red_add((uint32_t)tests);
    if (g.lane == 0) {
      hdr(H_NC) = (uint32_t)nC;
      counters()[KB_CNT_PAIR_TESTS] += tests;
    }
    if (g.any(overflow) && g.lane == 0) hdr(H_STATUS) |= KB_STATUS_CONTACT_OVERFLOW;
    g.sync();
  }

  // b2World::Step(dt, velIters, posIters)
  __device__ void worldStep() {
    if (g.lane == 0) {
      counters()[KB_CNT_SUBSTEPS] += 1ull;
      counters()[KB_CNT_CONTACTS] += (unsigned long long)hdr(H_NC);
    }
    collide();
    solve();
    if (L.enableToi) solveTOI();
  }

  __device__ void solveTOI();

  // ----------------------------------------------------------------------------- outputs
  __device__ void gather(const KernelArgs& a, int env) {
    g.sync();
    const int M = L.M, N = L.N;
    bool bad = false;
    for (int b = g.lane; b < L.B; b += LPE) {
      const float4 x = xf4(b);
      const float ang = pos4(b).z;
      bad |= !(isfinite(x.x) && isfinite(x.y) && isfinite(ang));
      float* o = b < M ? (a.obsObjects ? a.obsObjects + ((size_t)env * M + b) * 3 : nullptr)
                       : (a.obsKilobots ? a.obsKilobots + ((size_t)env * N + (b - M)) * 3 : nullptr);
      if (o) {
        o[0] = (float)((double)x.x / 25.0);
        o[1] = (float)((double)x.y / 25.0);
        o[2] = ang;
      }
    }
    if (g.any(bad) && g.lane == 0) hdr(H_STATUS) |= KB_STATUS_NONFINITE;
    if (a.obsLight) {
      const double* ls = lightState();
      for (int i = g.lane; i < L.L; i += LPE) a.obsLight[(size_t)env * L.L + i] = ls[i];
    }
    g.sync();
    if (g.lane == 0) {
      if (a.reward) a.reward[env] = __ldg(&a.scenes[a.envScene ? a.envScene[env] : 0].rewardConst);
      if (a.done) a.done[env] = 0;
      if (a.status) a.status[env] = (int32_t)hdr(H_STATUS);
    }
  }
};

}  // namespace kb
