// dynamics.cu -- batched robot_system update: one robot state per warp, one lane per body.
//
// Replaces, for B states at once, PinocchioRobotSystem.update_system and every getter the
// controller calls per tick (reference pnc/robot_system/pinocchio_robot_system.py:128-278):
// forward kinematics, frame placements / velocities / Jacobians / Jdot*qdot, centre of mass
// and its Jacobian (+ time derivative), CRBA mass matrix, gravity and Coriolis vectors.
//
// Formulation (B200-first, not pinocchio's body-frame recursions): every spatial quantity is
// expressed in WORLD axes about one reference point (the floating-base origin), so that
//   * a body's velocity / bias acceleration is a sum over its ANCESTORS of joint terms, and
//   * joint torques / composite inertias are sums over its SUBTREE of body terms.
// Bodies are numbered depth-first, so subtree(i) = lanes [i, i+subtree[i]) and both kinds of
// sum become masked loops over lanes through shared memory -- no per-level transforms on the
// way back up the tree.  Joint tables are staged once per CTA in shared memory.
#include "pnc_dev.cuh"
#include "pnc_launch.h"

namespace pnc {

#ifndef DYN_WARPS_N
#define DYN_WARPS_N 4  // warps (= states in flight) per CTA
#endif
constexpr int DYN_WARPS = DYN_WARPS_N;
#ifndef DYN_F64_CTAS
#define DYN_F64_CTAS 3  // resident CTAs per SM the fp64 instantiation is compiled for: 12 warps at 168 registers, no spills now that
                        // the phases hand their results on through shared memory (236 registers unconstrained; 8 warps: 0.89 ms,
                        // 10: 0.90, 12: 0.83 for 64 K Atlas states; shared memory stops at 12)
#endif
#ifndef DYN_F32_CTAS
#define DYN_F32_CTAS 4  // fp32 instantiation: 16 warps per SM at 128 registers
#endif
constexpr int XF_LD = 25;   // R(9) p(3) v(6) a(6) + pad (odd stride: conflict-free rows)
constexpr int SUB_LD = 19;  // m, h(3), Io(6), fc(6), P(3)
constexpr int SF_LD = 7;    // 6 + pad

struct DynSmemLayout {
  int per_warp_elems;  // in elements of the arithmetic type (8 bytes, or 4 in the fp32 mode)
  int off_xf, off_sub, off_S, off_F, off_S0, off_F0, off_fr, off_vloc, off_M;
};

__host__ __device__ inline DynSmemLayout dyn_layout(int nv) {
  DynSmemLayout L;
  int o = 0;
  L.off_sub = o; o += 32 * SUB_LD;
  L.off_S = o; o += 32 * SF_LD;
  L.off_F = o; o += 32 * SF_LD;
  L.off_S0 = o; o += 36;
  L.off_F0 = o; o += 36;
  L.off_fr = o; o += PNC_MAX_FRAMES * 4;  // frame world point (3) + pad
  L.off_vloc = o; o += 8;                 // free-flyer twist in the base frame [lin; ang] + pad
  // xf (phase 2..6) and the M staging tile (phase 7: packed lower triangle) share storage
  L.off_xf = o; L.off_M = o;
  int a = 32 * XF_LD, b = nv * (nv + 1) / 2;
  o += (a > b ? a : b);
  L.per_warp_elems = (o + 3) & ~3;
  return L;
}

size_t dyn_smem_bytes(int nv, size_t elem) {
  return sizeof(DevModel) + sizeof(DevFrames) + (size_t)DYN_WARPS * dyn_layout(nv).per_warp_elems * elem;
}

// ---- vector helpers on the kernel's arithmetic type T (double, or float in the optional fp32 mode) ----
// Table / input pointers may be double while T is float: every load converts to T first, so the
// T = double instantiation is operation for operation the original fp64 code.
template <typename T> struct V3T { T x, y, z; };
template <typename T> __device__ __forceinline__ V3T<T> operator+(V3T<T> a, V3T<T> b) { return V3T<T>{a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> __device__ __forceinline__ V3T<T> operator-(V3T<T> a, V3T<T> b) { return V3T<T>{a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> __device__ __forceinline__ V3T<T> operator*(T s, V3T<T> a) { return V3T<T>{s * a.x, s * a.y, s * a.z}; }
template <typename T> __device__ __forceinline__ V3T<T> cross(V3T<T> a, V3T<T> b) {
  return V3T<T>{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename T> __device__ __forceinline__ T dot(V3T<T> a, V3T<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T, typename A> __device__ __forceinline__ V3T<T> mulR(const A* R, V3T<T> a) {
  return V3T<T>{(T)R[0] * a.x + (T)R[1] * a.y + (T)R[2] * a.z, (T)R[3] * a.x + (T)R[4] * a.y + (T)R[5] * a.z,
                (T)R[6] * a.x + (T)R[7] * a.y + (T)R[8] * a.z};
}
template <typename T, typename A> __device__ __forceinline__ V3T<T> mulRt(const A* R, V3T<T> a) {
  return V3T<T>{(T)R[0] * a.x + (T)R[3] * a.y + (T)R[6] * a.z, (T)R[1] * a.x + (T)R[4] * a.y + (T)R[7] * a.z,
                (T)R[2] * a.x + (T)R[5] * a.y + (T)R[8] * a.z};
}
template <typename T, typename A, typename Bt> __device__ __forceinline__ void mul33(const A* a, const Bt* b, T* C) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      C[3 * i + j] = (T)a[3 * i] * (T)b[j] + (T)a[3 * i + 1] * (T)b[3 + j] + (T)a[3 * i + 2] * (T)b[6 + j];
}
template <typename T> __device__ __forceinline__ V3T<T> mulS(const T* S, V3T<T> a) {
  return V3T<T>{S[0] * a.x + S[1] * a.y + S[2] * a.z, S[1] * a.x + S[3] * a.y + S[4] * a.z,
                S[2] * a.x + S[4] * a.y + S[5] * a.z};
}
template <typename T, typename A> __device__ __forceinline__ V3T<T> ld3(const A* p) { return V3T<T>{(T)p[0], (T)p[1], (T)p[2]}; }
template <typename T, typename A> __device__ __forceinline__ void st3(A* p, V3T<T> a) { p[0] = (A)a.x; p[1] = (A)a.y; p[2] = (A)a.z; }
__device__ __forceinline__ void sincos_t(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ void sincos_t(float a, float* s, float* c) { sincosf(a, s, c); }
// scalar-last quaternion -> rotation, normalising like scipy Rotation.from_quat (quat_to_rot of pnc_dev.cuh on T)
template <typename T> __device__ __forceinline__ void quat_to_rot_t(const T* qq, T* R) {
  T nrm = sqrt(qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3]);
  T x = qq[0] / nrm, y = qq[1] / nrm, z = qq[2] / nrm, w = qq[3] / nrm;
  T x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
  T xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
  R[0] = x2 - y2 - z2 + w2; R[1] = 2 * (xy - zw);      R[2] = 2 * (xz + yw);
  R[3] = 2 * (xy + zw);     R[4] = -x2 + y2 - z2 + w2; R[5] = 2 * (yz - xw);
  R[6] = 2 * (xz - yw);     R[7] = 2 * (yz + xw);      R[8] = -x2 - y2 + z2 + w2;
}

// NV_ > 0 fixes the number of velocity variables at compile time (the shipped robots); 0 = read it
// from the model (any URDF).  Same code either way: index arithmetic on nv is what specialises.
// T = double: the fp64 path.  T = float: the optional fp32 mode (pnc_set_precision) -- the same formulation with every
// intermediate in fp32 (half the shared memory, the FP32 pipe); inputs are read and outputs stored as fp64, q is
// copied exactly, and the base position is carried in fp64 (it only enters sums at the outputs).
template <int NV_, typename T>
__global__ void __launch_bounds__(DYN_WARPS * 32, sizeof(T) == 4 ? DYN_F32_CTAS : DYN_F64_CTAS)
k_update_system(const DevModel* __restrict__ gm, const DevFrames* __restrict__ gfr, const int4* __restrict__ dof_tab,
                const int2* __restrict__ crba_pairs, int n_pairs, WsPtrs ws, int B,
                const double* __restrict__ base_pos, const double* __restrict__ base_quat,
                const double* __restrict__ base_lin_vel, const double* __restrict__ base_ang_vel,
                const double* __restrict__ joint_pos, const double* __restrict__ joint_vel, int b_cent) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DevModel* sm = reinterpret_cast<DevModel*>(smem_raw);
  DevFrames* sfr = reinterpret_cast<DevFrames*>(smem_raw + sizeof(DevModel));
  T* wbase = reinterpret_cast<T*>(smem_raw + sizeof(DevModel) + sizeof(DevFrames));
  using V3 = V3T<T>;
  using V3d = V3T<double>;
  auto mk = [](T x, T y, T z) { return V3{x, y, z}; };

  // stage the joint tables once per CTA
  {
    const int* src = reinterpret_cast<const int*>(gm);
    int* dst = reinterpret_cast<int*>(sm);
    for (int i = threadIdx.x; i < (int)(sizeof(DevModel) / 4); i += blockDim.x) dst[i] = src[i];
    const int* src2 = reinterpret_cast<const int*>(gfr);
    int* dst2 = reinterpret_cast<int*>(sfr);
    for (int i = threadIdx.x; i < (int)(sizeof(DevFrames) / 4); i += blockDim.x) dst2[i] = src2[i];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nb = sm->nb, nv = NV_ ? NV_ : sm->nv, na = sm->na, nfl = sm->nfl, nq = sm->nq, nfr = sfr->nfr;
  const DynSmemLayout L = dyn_layout(nv);
  T* W = wbase + (size_t)warp * L.per_warp_elems;
  T* xf = W + L.off_xf;
  T* sub = W + L.off_sub;
  T* Ss = W + L.off_S;
  T* Fs = W + L.off_F;
  T* S0 = W + L.off_S0;
  T* F0 = W + L.off_F0;
  T* frp = W + L.off_fr;
  T* vloc = W + L.off_vloc;
  T* Mst = W + L.off_M;

  const bool active = lane < nb;
  const int i = active ? lane : 0;
  // per-lane constants
  const int parent = sm->parent[i], jtype = sm->jtype[i], iv = sm->idx_v[i], nsub = sm->subtree[i],
            depth = sm->depth[i];
  const bool isff = active && jtype == PNC_JOINT_FREEFLYER;
  const T mass_i = (T)sm->mass[i];
  const T inv_mdyn = (T)(1.0 / sm->mass_dyn), inv_mtot = (T)(1.0 / sm->total_mass);
  const T GZ = (T)9.81;

  const int total_warps = gridDim.x * DYN_WARPS;
  for (int s = blockIdx.x * DYN_WARPS + warp; s < B; s += total_warps) {
    // Phases 0-2 hand their results to the later phases through shared memory only (xf, S, vloc): every phase
    // re-reads what it needs, so no register stays live from the tree propagation to the per-dof outputs.
    T qd = 0;
    V3d bpos = V3d{0, 0, 0};
    {
    // ---------------- phase 0/1: joint state and local transform -----------------------
    T Rl[9];  // joint frame in parent joint frame
    V3 pl;
    V3 vloc_lin = mk(0, 0, 0), vloc_ang = mk(0, 0, 0);  // free-flyer twist in the base frame
    if (isff) {
      const double qin[4] = {base_quat[4 * (size_t)s], base_quat[4 * (size_t)s + 1], base_quat[4 * (size_t)s + 2],
                             base_quat[4 * (size_t)s + 3]};
      const T qq[4] = {(T)qin[0], (T)qin[1], (T)qin[2], (T)qin[3]};
      T Rj[9];
      quat_to_rot_t(qq, Rj);
      mul33(sm->place_R[i], Rj, Rl);
      bpos = ld3<double>(base_pos + 3 * (size_t)s);
      pl = mk(0, 0, 0);  // reference point = base origin; bpos is added back on output
      // pinocchio_robot_system.py:149-162: world twist -> base frame
      vloc_lin = mulRt(Rj, ld3<T>(base_lin_vel + 3 * (size_t)s));
      vloc_ang = mulRt(Rj, ld3<T>(base_ang_vel + 3 * (size_t)s));
      double* qo = ws.q + (size_t)s * nq;
      qo[0] = bpos.x; qo[1] = bpos.y; qo[2] = bpos.z;
      qo[3] = qin[0]; qo[4] = qin[1]; qo[5] = qin[2]; qo[6] = qin[3];
      double* vo = ws.qdot + (size_t)s * nv;
      vo[0] = vloc_lin.x; vo[1] = vloc_lin.y; vo[2] = vloc_lin.z;
      vo[3] = vloc_ang.x; vo[4] = vloc_ang.y; vo[5] = vloc_ang.z;
      st3(vloc, vloc_lin); st3(vloc + 3, vloc_ang);
    } else if (active) {
      const double qjd = joint_pos[(size_t)s * na + iv - nfl], qdd_in = joint_vel[(size_t)s * na + iv - nfl];
      const T qj = (T)qjd;
      qd = (T)qdd_in;
      ws.q[(size_t)s * nq + (nfl ? iv + 1 : iv)] = qjd;
      ws.qdot[(size_t)s * nv + iv] = qdd_in;
      T sn, c;
      sincos_t(qj, &sn, &c);
      const T ax[3] = {(T)sm->axis[i][0], (T)sm->axis[i][1], (T)sm->axis[i][2]};
      T t = (T)1 - c, Rj[9];
      Rj[0] = c + t * ax[0] * ax[0];          Rj[1] = t * ax[0] * ax[1] - sn * ax[2]; Rj[2] = t * ax[0] * ax[2] + sn * ax[1];
      Rj[3] = t * ax[1] * ax[0] + sn * ax[2]; Rj[4] = c + t * ax[1] * ax[1];          Rj[5] = t * ax[1] * ax[2] - sn * ax[0];
      Rj[6] = t * ax[2] * ax[0] - sn * ax[1]; Rj[7] = t * ax[2] * ax[1] + sn * ax[0]; Rj[8] = c + t * ax[2] * ax[2];
      mul33(sm->place_R[i], Rj, Rl);
      pl = ld3<T>(sm->place_p[i]);
    }
    // the base position of this state, for every lane (lane 0 is the free-flyer when floating)
    if (nfl) {
      bpos.x = __shfl_sync(0xffffffffu, bpos.x, 0);
      bpos.y = __shfl_sync(0xffffffffu, bpos.y, 0);
      bpos.z = __shfl_sync(0xffffffffu, bpos.z, 0);
    }

    // ---------------- phase 2: propagate (R, p, v, a) down the tree, level by level ----
    T R[9];
    V3 p = mk(0, 0, 0), vl = mk(0, 0, 0), w = mk(0, 0, 0), al = mk(0, 0, 0), aw = mk(0, 0, 0);
    V3 Sl = mk(0, 0, 0), Sa = mk(0, 0, 0);  // own joint screw (revolute): motion [lin; ang] about the ref point
    for (int k = 0; k < 9; k++) R[k] = 0;
    const int maxdepth = sm->maxdepth;
    for (int d = 0; d <= maxdepth; d++) {
      if (active && depth == d) {
        V3 pp = mk(0, 0, 0), pvl = mk(0, 0, 0), pw = mk(0, 0, 0), pal = mk(0, 0, 0), paw = mk(0, 0, 0);
        if (parent >= 0) {
          const T* pr = xf + parent * XF_LD;
          mul33(pr, Rl, R);
          pp = ld3<T>(pr + 9);
          p = pp + mulR(pr, pl);
          pvl = ld3<T>(pr + 12); pw = ld3<T>(pr + 15); pal = ld3<T>(pr + 18); paw = ld3<T>(pr + 21);
        } else {
          for (int k = 0; k < 9; k++) R[k] = Rl[k];
          p = pl;
        }
        V3 vJl, vJa;
        if (isff) {
          vJl = mulR(R, vloc_lin);
          vJa = mulR(R, vloc_ang);  // reference point == base origin -> no p x w term
        } else {
          Sa = mulR(R, ld3<T>(sm->axis[i]));
          Sl = cross(p, Sa);
          vJl = qd * Sl;
          vJa = qd * Sa;
          T* sr = Ss + i * SF_LD;  // the joint's motion vector, for the per-dof outputs, the Jacobians and CRBA
          st3(sr, Sl); st3(sr + 3, Sa);
        }
        vl = pvl + vJl;
        w = pw + vJa;
        // a_i = a_parent + v_i x vJ_i   (spatial motion cross product; qddot = 0, no gravity)
        al = pal + cross(w, vJl) + cross(vl, vJa);
        aw = paw + cross(w, vJa);
        T* me = xf + i * XF_LD;
        for (int k = 0; k < 9; k++) me[k] = R[k];
        st3(me + 9, p); st3(me + 12, vl); st3(me + 15, w); st3(me + 18, al); st3(me + 21, aw);
      }
      __syncwarp();
    }
    }

    // ---------------- phase 3: world-frame body inertia, momentum, bias force ----------
    if (active) {
      const T* me = xf + i * XF_LD;
      T R[9];
      for (int k = 0; k < 9; k++) R[k] = me[k];
      const V3 p = ld3<T>(me + 9);
      V3 cw, h, P, fcl, fca;
      T Io[6];
      cw = p + mulR(R, ld3<T>(sm->com[i]));
      h = mass_i * cw;
      // Iw = R Ic R^T
      const T Ic[6] = {(T)sm->inertia[i][0], (T)sm->inertia[i][1], (T)sm->inertia[i][2],
                       (T)sm->inertia[i][3], (T)sm->inertia[i][4], (T)sm->inertia[i][5]};
      T Tm[9];
      // T = R * Ic (Ic symmetric)
      for (int r = 0; r < 3; r++) {
        Tm[3 * r + 0] = R[3 * r] * Ic[0] + R[3 * r + 1] * Ic[1] + R[3 * r + 2] * Ic[2];
        Tm[3 * r + 1] = R[3 * r] * Ic[1] + R[3 * r + 1] * Ic[3] + R[3 * r + 2] * Ic[4];
        Tm[3 * r + 2] = R[3 * r] * Ic[2] + R[3 * r + 1] * Ic[4] + R[3 * r + 2] * Ic[5];
      }
      T cc = dot(cw, cw);
      Io[0] = Tm[0] * R[0] + Tm[1] * R[1] + Tm[2] * R[2] + mass_i * (cc - cw.x * cw.x);
      Io[1] = Tm[0] * R[3] + Tm[1] * R[4] + Tm[2] * R[5] - mass_i * cw.x * cw.y;
      Io[2] = Tm[0] * R[6] + Tm[1] * R[7] + Tm[2] * R[8] - mass_i * cw.x * cw.z;
      Io[3] = Tm[3] * R[3] + Tm[4] * R[4] + Tm[5] * R[5] + mass_i * (cc - cw.y * cw.y);
      Io[4] = Tm[3] * R[6] + Tm[4] * R[7] + Tm[5] * R[8] - mass_i * cw.y * cw.z;
      Io[5] = Tm[6] * R[6] + Tm[7] * R[7] + Tm[8] * R[8] + mass_i * (cc - cw.z * cw.z);
      // momentum  (P, Lm) = I v ;  I a ;  bias force fc = I a + v x* (I v)
      const V3 vl = ld3<T>(me + 12), w = ld3<T>(me + 15), al = ld3<T>(me + 18), aw = ld3<T>(me + 21);
      P = mass_i * vl - cross(h, w);
      V3 Lm = cross(h, vl) + mulS(Io, w);
      V3 Ial = mass_i * al - cross(h, aw);
      V3 Iaa = cross(h, al) + mulS(Io, aw);
      fcl = Ial + cross(w, P);
      fca = Iaa + cross(w, Lm) + cross(vl, P);
      T* row = sub + i * SUB_LD;
      row[0] = mass_i; st3(row + 1, h);
      for (int k = 0; k < 6; k++) row[4 + k] = Io[k];
      st3(row + 10, fcl); st3(row + 13, fca); st3(row + 16, P);
    }
    __syncwarp();

    // ---------------- phase 4: subtree sums (composite inertia, bias force, momentum) ---
    // Reverse scan over the bodies with lanes = components: every body adds its (already complete)
    // row to its parent's.  Depth-first numbering puts children after their parent, so a row is
    // complete when its turn comes; lane k only ever touches column k, so no synchronisation is
    // needed inside the loop.  (Lanes = bodies would make the root lane walk all 31 rows while the
    // leaf lanes idle.)
    if (lane < SUB_LD) {
#pragma unroll 1
      for (int b = nb - 1; b > 0; b--) {
        const int pb = sm->parent[b];
        if (pb >= 0) sub[pb * SUB_LD + lane] += sub[b * SUB_LD + lane];
      }
    }
    __syncwarp();
    T acc[SUB_LD];
#pragma unroll
    for (int k = 0; k < SUB_LD; k++) acc[k] = active ? sub[i * SUB_LD + k] : (T)0;
    const T mS = acc[0];
    const V3 hS = mk(acc[1], acc[2], acc[3]);
    const T* IS = acc + 4;
    const V3 fSl = mk(acc[10], acc[11], acc[12]), fSa = mk(acc[13], acc[14], acc[15]);
    const V3 PS = mk(acc[16], acc[17], acc[18]);
    // whole-robot totals (lane 0's subtree is the whole robot on a single-rooted tree)
    T tot[7] = {0, 0, 0, 0, 0, 0, 0};  // h(3) P(3)... assembled below
    T tfcl[3] = {0, 0, 0};
    if (lane == 0) {
      tot[0] = acc[1]; tot[1] = acc[2]; tot[2] = acc[3]; tot[3] = acc[16]; tot[4] = acc[17]; tot[5] = acc[18];
      tfcl[0] = acc[10]; tfcl[1] = acc[11]; tfcl[2] = acc[12];
      for (int c = nsub; c < nb; c++) {  // forest (several bodies hang off the universe): add the other roots' subtrees
        if (sm->parent[c] >= 0) continue;
        const T* row = sub + c * SUB_LD;
        tot[0] += row[1]; tot[1] += row[2]; tot[2] += row[3]; tot[3] += row[16]; tot[4] += row[17]; tot[5] += row[18];
        tfcl[0] += row[10]; tfcl[1] += row[11]; tfcl[2] += row[12];
      }
      double* o = ws.com + 3 * (size_t)s;
      o[0] = bpos.x + tot[0] * inv_mdyn; o[1] = bpos.y + tot[1] * inv_mdyn; o[2] = bpos.z + tot[2] * inv_mdyn;
      o = ws.vcom + 3 * (size_t)s;
      o[0] = tot[3] * inv_mdyn; o[1] = tot[4] * inv_mdyn; o[2] = tot[5] * inv_mdyn;
      o = ws.jcomdot_qdot + 3 * (size_t)s;  // (dAg_lin . qdot) / total_mass, basic_task.py:133-135
      o[0] = tfcl[0] * inv_mtot; o[1] = tfcl[1] * inv_mtot; o[2] = tfcl[2] * inv_mtot;
    }
    // centroidal quantities (pinocchio_robot_system.py:181-194, pin.ccrba): optional
    V3 crel = mk(0, 0, 0);  // centre of mass relative to the reference point
    if (b_cent) {
      T mt = 0, Iot[6] = {0, 0, 0, 0, 0, 0};
      if (lane == 0) {
        mt = acc[0];
        for (int k = 0; k < 6; k++) Iot[k] = acc[4 + k];
        for (int c = nsub; c < nb; c++) {
          if (sm->parent[c] >= 0) continue;
          const T* row = sub + c * SUB_LD;
          mt += row[0];
          for (int k = 0; k < 6; k++) Iot[k] += row[4 + k];
        }
        crel = mk(tot[0] / mt, tot[1] / mt, tot[2] / mt);
        // Ig: rotational inertia about the CoM (world axes) and m I3, off-diagonal blocks dropped (:192-194)
        double* ig = ws.cent_ig + 36 * (size_t)s;
        for (int k = 0; k < 36; k++) ig[k] = 0.0;
        const T cc = dot(crel, crel);
        const T cx[3] = {crel.x, crel.y, crel.z};
        const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) {
            ig[6 * r + c] = Iot[sym[r][c]] - mt * ((r == c ? cc : (T)0) - cx[r] * cx[c]);
            ig[6 * (3 + r) + 3 + c] = r == c ? mt : (T)0;
          }
      }
      crel.x = __shfl_sync(0xffffffffu, crel.x, 0);
      crel.y = __shfl_sync(0xffffffffu, crel.y, 0);
      crel.z = __shfl_sync(0xffffffffu, crel.z, 0);
    }
    T hgp[6] = {0, 0, 0, 0, 0, 0};

    // ---------------- phase 5: per-dof outputs: gravity, Coriolis, CoM Jacobians, F = Ic S
    if (active) {
      const int ndof = isff ? 6 : 1;
      const T* me = xf + i * XF_LD;
      const V3 vl = ld3<T>(me + 12), w = ld3<T>(me + 15);
      double* grav = ws.grav + (size_t)s * nv;
      double* cori = ws.cori + (size_t)s * nv;
      double* jc = ws.jcom + (size_t)s * 3 * nv;
      double* jcd = ws.jcomdot + (size_t)s * 3 * nv;
      const V3 gvec = mk(0, 0, GZ);
      const V3 fgl = mS * gvec, fga = cross(hS, gvec);
      for (int d = 0; d < ndof; d++) {
        V3 sl, sa;
        if (isff) {
          V3 col = mk(me[d % 3], me[3 + d % 3], me[6 + d % 3]);  // column d % 3 of R
          if (d < 3) { sl = col; sa = mk(0, 0, 0); }
          else { sa = col; sl = cross(ld3<T>(me + 9), col); }
          T* s0 = S0 + d * 6;
          st3(s0, sl); st3(s0 + 3, sa);
        } else {
          const T* sr = Ss + i * SF_LD;  // written in phase 2
          sl = ld3<T>(sr); sa = ld3<T>(sr + 3);
        }
        grav[iv + d] = dot(sl, fgl) + dot(sa, fga);
        cori[iv + d] = dot(sl, fSl) + dot(sa, fSa);
        V3 jcol = inv_mdyn * (mS * sl + cross(sa, hS));
        jc[iv + d] = jcol.x; jc[nv + iv + d] = jcol.y; jc[2 * nv + iv + d] = jcol.z;
        // dS = v_i x S ;  dAg_lin col = -Pdot_sub x S_ang + m dS_lin - h x dS_ang
        V3 dsl = cross(w, sl) + cross(vl, sa), dsa = cross(w, sa);
        V3 dcol = inv_mtot * (mS * dsl - cross(PS, sa) - cross(hS, dsa));
        jcd[iv + d] = dcol.x; jcd[nv + iv + d] = dcol.y; jcd[2 * nv + iv + d] = dcol.z;
        // F = Ic_sub S : lin = m S_lin - h x S_ang ; ang = h x S_lin + Io S_ang
        V3 Fl = mS * sl - cross(hS, sa);
        V3 Fa = cross(hS, sl) + mulS(IS, sa);
        T* fr = isff ? (F0 + d * 6) : (Fs + i * SF_LD);
        st3(fr, Fl); st3(fr + 3, Fa);
        if (b_cent) {  // Ag column = [moment about the CoM; force], rows [angular; linear]
          const V3 ac = Fa - cross(crel, Fl);
          double* ag = ws.cent_ag + (size_t)s * 6 * nv;
          ag[iv + d] = ac.x; ag[nv + iv + d] = ac.y; ag[2 * nv + iv + d] = ac.z;
          ag[3 * nv + iv + d] = Fl.x; ag[4 * nv + iv + d] = Fl.y; ag[5 * nv + iv + d] = Fl.z;
          const T qdd = isff ? vloc[d] : qd;
          hgp[0] += ac.x * qdd; hgp[1] += ac.y * qdd; hgp[2] += ac.z * qdd;
          hgp[3] += Fl.x * qdd; hgp[4] += Fl.y * qdd; hgp[5] += Fl.z * qdd;
        }
      }
    }
    if (b_cent) {  // hg = Ag qdot
#pragma unroll
      for (int k = 0; k < 6; k++) {
        T v = hgp[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) ws.cent_hg[6 * (size_t)s + k] = v;
      }
    }
    __syncwarp();

    // ---------------- phase 6: frames (placement, velocity, Jdot*qdot, Jacobian columns) --
    if (lane < nfr) {
      const int f = lane, b = sfr->body[f];
      T Rw[9];
      V3 pf, fvl = mk(0, 0, 0), fw = mk(0, 0, 0), fal = mk(0, 0, 0), faw = mk(0, 0, 0);
      if (b >= 0) {
        const T* br = xf + b * XF_LD;
        mul33(br, sfr->R[f], Rw);
        pf = ld3<T>(br + 9) + mulR(br, ld3<T>(sfr->p[f]));
        V3 bvl = ld3<T>(br + 12), bw = ld3<T>(br + 15), bal = ld3<T>(br + 18), baw = ld3<T>(br + 21);
        fw = bw;
        fvl = bvl + cross(bw, pf);
        faw = baw;
        fal = bal + cross(baw, pf) + cross(bw, fvl);  // classical acceleration: + w x v
      } else {
        for (int k = 0; k < 9; k++) Rw[k] = (T)sfr->R[f][k];
        pf = mk((T)(sfr->p[f][0] - bpos.x), (T)(sfr->p[f][1] - bpos.y), (T)(sfr->p[f][2] - bpos.z));
      }
      st3(frp + 4 * f, pf);
      double* iso = ws.fr_iso + ((size_t)s * nfr + f) * 12;
      for (int k = 0; k < 9; k++) iso[k] = Rw[k];
      iso[9] = pf.x + bpos.x; iso[10] = pf.y + bpos.y; iso[11] = pf.z + bpos.z;
      double* fv = ws.fr_vel + ((size_t)s * nfr + f) * 6;
      st3(fv, fw); st3(fv + 3, fvl);
      double* fa = ws.fr_jdqd + ((size_t)s * nfr + f) * 6;
      st3(fa, faw); st3(fa + 3, fal);
    }
    __syncwarp();
    // Jacobian columns, one (frame, dof) item at a time dealt round-robin to the lanes (with lanes = bodies the
    // root lane wrote its six columns while the others waited): the dof's motion vector is in shared memory
    // (phase 5), its body's subtree range comes from a table built with the model; consecutive lanes write
    // consecutive columns
    {
      const int total = nfr * nv;
      for (int e = lane; e < total; e += 32) {
        const int f = e / nv, c = e - f * nv;
        const int4 dt = __ldg(dof_tab + c);  // x: offset of S, [y, z): bodies of the dof's subtree
        const int b = sfr->body[f];
        const bool on_path = b >= dt.y && b < dt.z;
        const T* Sd = W + dt.x;
        const V3 sl = ld3<T>(Sd), sa = ld3<T>(Sd + 3);
        const V3 pf = ld3<T>(frp + 4 * f);
        double* J = ws.fr_jac + ((size_t)s * nfr + f) * 6 * nv;
        V3 lin = mk(0, 0, 0), ang = mk(0, 0, 0);
        if (on_path) { ang = sa; lin = sl + cross(sa, pf); }
        // rows [ang; lin], pinocchio_robot_system.py:259-261
        J[c] = ang.x; J[nv + c] = ang.y; J[2 * nv + c] = ang.z;
        J[3 * nv + c] = lin.x; J[4 * nv + c] = lin.y; J[5 * nv + c] = lin.z;
      }
    }
    __syncwarp();  // everyone is done with xf before the M tile overwrites it

    // ---------------- phase 7: CRBA rows into a shared tile, then coalesced store -------
    for (int k = lane; k < nv * (nv + 1) / 2; k += 32) Mst[k] = (T)0;
    __syncwarp();
    // entries between each dof and its ancestor-or-self dofs, M[a, b] = F_a . S_b: the pairs come from a table
    // built with the model (build_crba_pairs) and are dealt round-robin to the lanes -- with lanes = bodies the
    // root lane walked 31 candidate ancestors and 36 free-flyer products while the leaves idled
    for (int e = lane; e < n_pairs; e += 32) {
      const int2 pr = __ldg(crba_pairs + e);
      const T* Fi = W + (pr.x & 0xffff);
      const T* Sj = W + (pr.x >> 16);
      const int ra = pr.y & 0xff, cb = pr.y >> 8;
      T val = Fi[0] * Sj[0] + Fi[1] * Sj[1] + Fi[2] * Sj[2] + Fi[3] * Sj[3] + Fi[4] * Sj[4] + Fi[5] * Sj[5];
      Mst[ra * (ra + 1) / 2 + cb] = val;  // ra >= cb: a dof's ancestors come first in the depth-first numbering
    }
    __syncwarp();
    {
      double* Mo = ws.M + (size_t)s * nv * nv;
      for (int k = lane; k < nv * nv; k += 32) {  // both triangles from the packed one: M is exactly symmetric
        const int r = k / nv, c = k - r * nv;
        const int hi = r > c ? r : c, lo = r > c ? c : r;
        Mo[k] = Mst[hi * (hi + 1) / 2 + lo];
      }
    }
    __syncwarp();
  }
}

void build_dof_table(const DevModel& h, int4* tab) {
  const DynSmemLayout L = dyn_layout(h.nv);
  for (int i = 0; i < h.nb; i++) {
    const bool iff = h.jtype[i] == PNC_JOINT_FREEFLYER;
    for (int d = 0; d < (iff ? 6 : 1); d++)
      tab[h.idx_v[i] + d] = make_int4(iff ? L.off_S0 + d * 6 : L.off_S + i * SF_LD, i, i + h.subtree[i], 0);
  }
}

int build_crba_pairs(const DevModel& h, int2* pairs, int max_pairs) {
  const DynSmemLayout L = dyn_layout(h.nv);
  int n = 0;
  for (int i = 0; i < h.nb; i++) {
    const bool iff = h.jtype[i] == PNC_JOINT_FREEFLYER;
    for (int j = 0; j <= i; j++) {
      if (!(i < j + h.subtree[j])) continue;  // j is not an ancestor of i (nor i itself)
      const bool jff = h.jtype[j] == PNC_JOINT_FREEFLYER;
      for (int di = 0; di < (iff ? 6 : 1); di++)
        for (int dj = 0; dj < (jff ? 6 : 1); dj++) {
          if (i == j && dj > di) continue;  // free-flyer block: lower triangle, mirrored by the kernel
          const int offF = iff ? L.off_F0 + di * 6 : L.off_F + i * SF_LD;
          const int offS = jff ? L.off_S0 + dj * 6 : L.off_S + j * SF_LD;
          if (n < max_pairs) pairs[n] = make_int2(offF | (offS << 16), (h.idx_v[i] + di) | ((h.idx_v[j] + dj) << 8));
          n++;
        }
    }
  }
  return n <= max_pairs ? n : -1;
}

void launch_update_system(const DevModel* dm, const DevFrames* dfr, const int4* dof_tab, const int2* crba_pairs, int n_pairs,
                          const WsPtrs& ws, int B, int nv, const double* base_pos, const double* base_quat, const double* base_lin_vel,
                          const double* base_ang_vel, const double* joint_pos, const double* joint_vel,
                          int num_sms, cudaStream_t stream, int b_cent, int fp32) {
  size_t smem = dyn_smem_bytes(nv, fp32 ? sizeof(float) : sizeof(double));
  const int need = (B + DYN_WARPS - 1) / DYN_WARPS;
  auto go = [&](auto kernel) {
    static SmemOptIn optin;
    optin.ensure(kernel, smem);
    // persistent grid: as many CTAs as are resident at once (registers and shared memory both count).  Cached per
    // (kernel, device, shared-memory size): every instantiation has the same pointer TYPE, so a plain static would be shared.
    struct Entry { const void* fn; int dev; size_t smem; int ctas; };
    static Entry table[64];
    static int used = 0;
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    const void* fn = reinterpret_cast<const void*>(kernel);
    int ctas_per_sm = 0;
    {
      std::lock_guard<std::mutex> lock(mu);
      for (int i = 0; i < used; i++)
        if (table[i].fn == fn && table[i].dev == dev && table[i].smem == smem) ctas_per_sm = table[i].ctas;
      if (ctas_per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kernel, DYN_WARPS * 32, smem) != cudaSuccess || ctas_per_sm < 1)
          ctas_per_sm = 1;
        if (used < 64) table[used++] = Entry{fn, dev, smem, ctas_per_sm};
      }
    }
    int grid = num_sms * ctas_per_sm;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kernel<<<grid, DYN_WARPS * 32, smem, stream>>>(dm, dfr, dof_tab, crba_pairs, n_pairs, ws, B, base_pos, base_quat, base_lin_vel,
                                                     base_ang_vel, joint_pos, joint_vel, b_cent);
  };
  if (fp32) {
    switch (nv) {
      case 36: go(k_update_system<36, float>); break;
      case 33: go(k_update_system<33, float>); break;
      case 32: go(k_update_system<32, float>); break;
      case 20: go(k_update_system<20, float>); break;
      case 18: go(k_update_system<18, float>); break;
      default: go(k_update_system<0, float>); break;
    }
  } else {
    switch (nv) {  // Atlas, Draco3, Valkyrie, Draco3 lower body, Laikago; anything else reads nv at run time
      case 36: go(k_update_system<36, double>); break;
      case 33: go(k_update_system<33, double>); break;
      case 32: go(k_update_system<32, double>); break;
      case 20: go(k_update_system<20, double>); break;
      case 18: go(k_update_system<18, double>); break;
      default: go(k_update_system<0, double>); break;
    }
  }
  g_launch_count++;
}

}  // namespace pnc
