// TEST INFRASTRUCTURE ONLY (see simt_emu.h): compiles the kernel source of the
// product with g++ under a lock-step SIMT emulation and exposes a few entry
// points so the CPU-only test tier can exercise the device code paths
// (RHS, LU, Jacobian, BDF stepper) without a GPU.  Never linked into
// libchemtools_b200.so and never imported by the chemtools_b200 package.
#include "simt_emu.h"
struct double2 { double x, y; };
#include "../../chemtools_b200/csrc/kernels.cuh"
#include "../../chemtools_b200/csrc/mechanism.hpp"

#include <memory>
#include <string>

using namespace ctb;

namespace {
std::string g_err;
struct EmuMech { MechanismDef def; CompiledMech C; MechDesc d; };

MechDesc desc_of(const CompiledMech& C) {
  MechDesc d{};
  d.K = C.K; d.R = C.R; d.nf = C.n_fall; d.n3 = C.n_3b; d.KP = C.KP; d.nc = C.n_const; d.nx = C.n_exp_end; d.Rgas = C.gas_constant; d.p_ref = C.p_ref; d.kc_Tmin = C.kc_fast_Tmin;
  const auto& o = C.off;
  d.o_W = o.W; d.o_rW = o.rW; d.o_Tmid = o.Tmid; d.o_nasa = o.nasa; d.o_A = o.A; d.o_b = o.b; d.o_E = o.E; d.o_A0 = o.A0;
  d.o_b0 = o.b0; d.o_E0 = o.E0; d.o_ta = o.troe_a; d.o_rt3 = o.troe_rt3; d.o_rt1 = o.troe_rt1; d.o_t2 = o.troe_t2;
  d.o_eff_val = o.eff_val; d.o_eff_ptr = o.eff_ptr; d.o_sp_ptr = o.sp_ptr; d.o_sp_rxn = o.sp_rxn; d.o_sp_nu = o.sp_nu;
  d.o_eff_o = o.eff_o; d.o_ftype = o.ftype; d.o_rxo = o.rxo; d.o_ent = o.ent; d.o_pstart = o.pstart; d.o_pnplus = o.pnplus;
  d.o_lpp = o.lpp; d.o_spp = o.spp; d.n_pieces = o.n_pieces;
  d.blob_bytes = o.total;
  static_assert(sizeof(MechDesc) <= CT_GANG_OFF && CT_POOL_OFF + 16 <= CT_HDR_BYTES, "MechDesc must fit the shared header");
  return d;
}
}  // namespace

namespace {
int g_static = 0;  // emu_use_static_shapes: run the compile-time-shape instantiations where the mechanism matches
template <class SH>
bool shape_matches(const MechDesc& d) {
  const int* w = SH::words();
  const int* a = reinterpret_cast<const int*>(&d);
  const int* b = reinterpret_cast<const int*>(&d.o_W);
  for (int i = 0; i < CT_DESC_INTS_A; ++i) if (a[i] != w[i]) return false;
  for (int i = 0; i < CT_DESC_INTS_B; ++i) if (b[i] != w[CT_DESC_INTS_A + i]) return false;
  return true;
}
int shape_of(const MechDesc& d) {
  if (!g_static) return 0;
  return shape_matches<ShapeGri30>(d) ? 1 : shape_matches<ShapeGri211>(d) ? 2 : shape_matches<ShapeGri12>(d) ? 3 : 0;
}
template <class... A>
void launch_integrate(const MechDesc& d, unsigned block, size_t smem, A... args) {
  switch (shape_of(d)) {
    case 1: emu_launch(k_integrate<256, ShapeGri30>, 1, block, smem, args...); break;
    case 2: emu_launch(k_integrate<256, ShapeGri211>, 1, block, smem, args...); break;
    case 3: emu_launch(k_integrate<256, ShapeGri12>, 1, block, smem, args...); break;
    default: emu_launch(k_integrate<256, ShapeRt>, 1, block, smem, args...);
  }
}
}  // namespace

extern "C" {

static double* g_stats_ex = nullptr;  // optional [n x 6] buffer for the next emu_integrate call
void emu_set_stats_ex_buffer(double* p) { g_stats_ex = p; }
void emu_use_static_shapes(int on) { g_static = on; }
int emu_static_shape(void* m) { const int keep = g_static; g_static = 1; const int r = shape_of(((EmuMech*)m)->d); g_static = keep; return r; }

const char* emu_last_error() { return g_err.c_str(); }

void* emu_mech_load(const char* path, const char* constants) {
  try {
    std::unique_ptr<EmuMech> m(new EmuMech);
    std::string p(path);
    bool ctm = p.size() > 4 && p.compare(p.size() - 4, 4, ".ctm") == 0;
    m->def = ctm ? read_ctm(p) : parse_cti(p, nullptr);
    m->C = compile_mechanism(m->def, constants ? constants : "");
    m->d = desc_of(m->C);
    return m.release();
  } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void emu_mech_free(void* m) { delete (EmuMech*)m; }

static void run_init(EmuMech* m, int rt, long long n, const double* T0, const double* p0, const double* Y0, double* state, double* rho0, int N) {
  emu_launch(k_init, 1, 32, 0, m->d, (const unsigned char*)m->C.blob.data(), rt, n, T0, p0, Y0, state, rho0, N);
}

int emu_rhs(void* mv, int rt, long long n, const double* T0, const double* p0, const double* Y0, const double* states,
            double* ydot, double* wdot, double* qf, double* qr) {
  EmuMech* m = (EmuMech*)mv;
  const int N = rt_has_energy(rt) ? m->C.K + 1 : m->C.K;
  std::vector<double> st0((size_t)n * N), rho0(n);
  run_init(m, rt, n, T0, p0, Y0, st0.data(), rho0.data(), N);
  BatchArgs a{};
  a.rt = rt; a.N = N; a.n = n; a.T0 = T0; a.p0 = p0; a.rho0 = rho0.data(); a.rtol = 1e-8; a.atol = 1e-14;
  const WarpLayout L = make_layout(N, m->d.R);
  const size_t smem = CT_HDR_BYTES + m->d.blob_bytes + (size_t)L.doubles * 8;
  auto kern = shape_of(m->d) == 1 ? k_rhs_t<256, ShapeGri30> : shape_of(m->d) == 2 ? k_rhs_t<256, ShapeGri211> : shape_of(m->d) == 3 ? k_rhs_t<256, ShapeGri12> : k_rhs;
  emu_launch(kern, 1, 32, smem, m->d, (const unsigned char*)m->C.blob.data(), a, states, (const double*)nullptr, ydot, wdot, qf, qr,
             (const int*)m->C.perm.data());
  return 0;
}

int emu_jacobian(void* mv, int rt, int mode, long long n, const double* T0, const double* p0, const double* Y0,
                 const double* states, double* J) {
  EmuMech* m = (EmuMech*)mv;
  const int N = rt_has_energy(rt) ? m->C.K + 1 : m->C.K;
  std::vector<double> st0((size_t)n * N), rho0(n), scratch(warp_scratch_doubles(N, m->d.R));
  run_init(m, rt, n, T0, p0, Y0, st0.data(), rho0.data(), N);
  BatchArgs a{};
  a.rt = rt; a.N = N; a.n = n; a.T0 = T0; a.p0 = p0; a.rho0 = rho0.data(); a.rtol = 1e-8; a.atol = 1e-14; a.jac_mode = mode; a.jac_clip = 1.0;
  a.scratch = scratch.data();
  const WarpLayout L = make_layout(N, m->d.R);
  const size_t smem = CT_HDR_BYTES + m->d.blob_bytes + (size_t)L.doubles * 8;
  emu_launch(k_jacobian, 1, 32, smem, m->d, (const unsigned char*)m->C.blob.data(), a, states, (const double*)nullptr, 1e-8, J);
  return 0;
}

int emu_lu_solve(int N, long long n, const double* A, const double* b, double* x, int* info, int linsol) {
  const WarpLayout L = make_layout(N, 4);
  emu_launch(k_lu_solve, 1, 32, (size_t)L.doubles * 8, N, n, A, b, x, info, 1, linsol);
  return 0;
}

int emu_integrate(void* mv, int rt, long long n, const double* T0, const double* p0, const double* Y0, double rtol, double atol,
                  double t_end, long long mxstep, int jac_mode, int ign_mode, double ign_value, int n_out, int n_calls,
                  int nt, const double* tab_t, const double* tab_T, const double* tab_p,
                  double* traj, double* state, double* tau, long long* stats, int* status, int warps, int linsol, double jac_clip, int pool_n,
                  int slice_steps, const double* opts /* 8 or null: h_init h_min h_max nls_coef max_nef max_cor max_ncf max_order */) {
  EmuMech* m = (EmuMech*)mv;
  const int N = rt_has_energy(rt) ? m->C.K + 1 : m->C.K;
  std::vector<double> rho0(n), tcur(n);
  run_init(m, rt, n, T0, p0, Y0, state, rho0.data(), N);
  if (warps < 1) warps = 1;
  const bool sliced = slice_steps > 0 && pool_n > 0;
  std::vector<double> scratch((size_t)(sliced ? std::max<long long>(n, warps) : warps) * warp_scratch_doubles(N, m->d.R));
  unsigned long long counter = 0;
  const int cap = (int)(2 * n);
  std::vector<int> ring(cap), yielded(n);
  std::vector<unsigned long long> ring_seq(cap), qctl(4);
  std::vector<double> slice_save;
  BatchArgs a{};
  a.rt = rt; a.N = N; a.jac_mode = jac_mode; a.max_order = 5; a.ign_mode = ign_mode; a.fresh = 1; a.n_out = n_out; a.n = n;
  a.mxstep = mxstep; a.T0 = T0; a.p0 = p0; a.rho0 = rho0.data(); a.rtol = rtol; a.atol = atol; a.t_target = t_end; a.t_stop = t_end;
  a.nt = nt; a.tab_t = tab_t; a.tab_T = tab_T; a.tab_p = tab_p;
  a.ign_value = ign_value; a.state = state; a.traj = traj; a.tau = tau; a.tcur = tcur.data(); a.stats = stats; a.status = status;
  a.counter = &counter; a.scratch = scratch.data(); a.linsol = linsol; a.jac_clip = jac_clip; a.gang = 1; a.pool_n = pool_n;
  a.save_stride = CT_SAVE_DOUBLES;
  a.stats_ex = g_stats_ex; g_stats_ex = nullptr;
  if (opts) {
    a.opts.h_init = opts[0]; a.opts.h_min = opts[1]; a.opts.h_max_inv = opts[2] > 0 ? 1.0 / opts[2] : 0.0; a.opts.nls_coef = opts[3];
    a.opts.max_nef = (int)opts[4]; a.opts.max_cor = (int)opts[5]; a.opts.max_ncf = (int)opts[6];
    if (opts[7] > 0) a.max_order = (int)opts[7];
  }
  const WarpLayout L = make_layout(N, m->d.R, pool_n == 0);
  const size_t smem = ((CT_HDR_BYTES + m->d.blob_bytes + (size_t)warps * L.doubles * 8 + 15) & ~(size_t)15) + (size_t)pool_n * pool_slab_bytes(N);
  std::vector<double> save;
  if (sliced) {
    slice_save.assign((size_t)n * a.save_stride, 0.0);
    a.slice_steps = slice_steps; a.ring_cap = cap; a.ring = ring.data(); a.ring_seq = ring_seq.data(); a.qctl = qctl.data();
    a.yielded = yielded.data(); a.save = slice_save.data();
  }
  auto ring_init = [&]() { if (sliced) emu_launch(k_ring_init, (cap + 255) / 256, 256, 0, (int)n, cap, (const int*)nullptr, a.ring, a.ring_seq, a.qctl, a.yielded); };
  for (long long i = 0; i < n; ++i) status[i] = 0;
  if (n_calls <= 1) {
    ring_init();
    launch_integrate(m->d, 32 * warps, smem, m->d, (const unsigned char*)m->C.blob.data(), a);
  } else {
    // repeated ct_batch_integrate(t) calls with increasing targets (resumable state)
    save.assign((size_t)n * a.save_stride, 0.0);
    a.save = save.data(); a.user_save = 1;
    a.n_out = 0; a.traj = nullptr;
    for (int c = 0; c < n_calls; ++c) {
      a.t_target = (c + 1 == n_calls) ? t_end : (c + 1) * (t_end / n_calls);
      counter = 0;
      ring_init();
      launch_integrate(m->d, 32 * warps, smem, m->d, (const unsigned char*)m->C.blob.data(), a);
      a.fresh = 0;
      if (traj) for (long long i = 0; i < n; ++i) for (int k = 0; k < N; ++k) traj[((size_t)i * n_calls + c) * N + k] = state[i * N + k];
    }
  }
  return 0;
}

int emu_roberts(int n_out, const double* t_out, double rtol, const double* atol3, double* y_out, long long* stats8) {
  MechDesc d{};
  double state[3] = {1.0, 0.0, 0.0}, tau = 0, tcur = 0;
  int status = 0;
  unsigned long long counter = 0;
  std::vector<double> scratch(warp_scratch_doubles(3, 0) + 16);
  BatchArgs a{};
  a.rt = RT_ROBERTS; a.N = 3; a.jac_mode = JAC_ANALYTIC; a.max_order = 5; a.fresh = 1; a.n_out = n_out; a.n = 1; a.mxstep = 500;
  a.rtol = rtol; a.atol = 0.0; a.atol_vec = atol3; a.t_target = 1e300; a.t_stop = 1e300; a.state = state; a.traj = y_out;
  a.tau = &tau; a.tcur = &tcur; a.stats = stats8; a.status = &status; a.counter = &counter; a.scratch = scratch.data();
  a.stats_ex = g_stats_ex; g_stats_ex = nullptr;
  a.out_times = t_out; a.linsol = LINSOL_LU;
  const WarpLayout L = make_layout(3, 0);
  emu_launch(k_integrate<256, ShapeRt>, 1, 32, CT_HDR_BYTES + (size_t)L.doubles * 8, d, (const unsigned char*)nullptr, a);
  return status;
}

}  // extern "C"
