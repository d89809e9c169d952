// C ABI (include/chemtools_b200.h) over the CUDA kernels in kernels.cuh.
// Host side is C++14; there is no CPU compute path in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/chemtools_b200.h"
#include "kernels.cuh"
#include "h5_writer.hpp"
#include "input_parser.hpp"
#include "mechanism.hpp"

#include <sys/stat.h>

using namespace ctb;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CT_CUDA(call)                                                                                  \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess)                                                                             \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? CT_ERR_NO_DEVICE : CT_ERR_CUDA, \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                                 \
  } while (0)

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(CT_ERR_NO_DEVICE, "no CUDA device available: chemtools_b200 has no CPU fallback");
  }
  return 0;
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  cudaError_t alloc(size_t count) {
    if (count == n && p) return cudaSuccess;
    release();
    n = count;
    return cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
  }
};

MechDesc make_desc(const CompiledMech& C) {
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
  static_assert(offsetof(MechDesc, Rgas) == 4 * CT_DESC_INTS_A && offsetof(MechDesc, o_W) == 4 * CT_DESC_INTS_A + 24 &&
                offsetof(MechDesc, blob_bytes) == 4 * CT_DESC_INTS_A + 24 + 4 * (CT_DESC_INTS_B - 1), "MechDesc word layout (CT_STATIC_SHAPE)");
  return d;
}

// which compile-time shape (kernels.cuh: CT_STATIC_SHAPE, mech_shapes.inc) a mechanism matches word for word: 1 GRI-3.0,
// 2 GRI-2.11, 3 GRI-1.2, 0 none (run-time shape)
template <class SH>
bool shape_matches(const MechDesc& d) {
  const int* w = SH::words();
  const int* a = reinterpret_cast<const int*>(&d);
  const int* b = reinterpret_cast<const int*>(&d.o_W);
  for (int i = 0; i < CT_DESC_INTS_A; ++i) if (a[i] != w[i]) return false;
  for (int i = 0; i < CT_DESC_INTS_B; ++i) if (b[i] != w[CT_DESC_INTS_A + i]) return false;
  return true;
}
int static_shape_of(const MechDesc& d) {
  if (shape_matches<ShapeGri30>(d)) return 1;
  if (shape_matches<ShapeGri211>(d)) return 2;
  if (shape_matches<ShapeGri12>(d)) return 3;
  return 0;
}

struct DeviceInfo { int sms = 0; int max_smem_optin = 0; };
int device_info(DeviceInfo& di) {
  int dev = 0;
  CT_CUDA(cudaGetDevice(&dev));
  CT_CUDA(cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev));
  CT_CUDA(cudaDeviceGetAttribute(&di.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return 0;
}

}  // namespace

struct ct_mech {
  MechanismDef def;
  CompiledMech C;
  MechDesc desc{};
  int shape = 0;  // static_shape_of(desc)
  std::map<int, std::pair<unsigned char*, int*>> dev;  // device ordinal -> (blob, perm)
  std::mutex mu;  // batches on several devices may be driven from several host threads (ct_ensemble)
  ~ct_mech() {
    for (auto& kv : dev) { cudaFree(kv.second.first); cudaFree(kv.second.second); }
  }
  int on_device(const unsigned char** blob, const int** perm) {
    int d = 0;
    CT_CUDA(cudaGetDevice(&d));
    std::lock_guard<std::mutex> lock(mu);
    auto it = dev.find(d);
    if (it == dev.end()) {
      unsigned char* b = nullptr;
      int* p = nullptr;
      CT_CUDA(cudaMalloc(&b, C.blob.size()));
      CT_CUDA(cudaMemcpy(b, C.blob.data(), C.blob.size(), cudaMemcpyHostToDevice));
      CT_CUDA(cudaMalloc(&p, C.perm.size() * sizeof(int)));
      CT_CUDA(cudaMemcpy(p, C.perm.data(), C.perm.size() * sizeof(int), cudaMemcpyHostToDevice));
      it = dev.emplace(d, std::make_pair(b, p)).first;
    }
    *blob = it->second.first;
    if (perm) *perm = it->second.second;
    return 0;
  }
};

struct ct_batch {
  ct_mech* mech = nullptr;
  int rt = 0, N = 0, K = 0;
  int64_t n = 0;
  double rtol = 1e-8, atol = 1e-14, t_stop = 1e-3;
  int64_t mxstep = 500;
  StepOpts opts{};      // solver option surface; zero members = CVODE defaults
  double t_start = 0.0; // integration start of a fresh batch (ct_batch_reinit)
  int max_order = 5, jac_mode = CT_JAC_ANALYTIC, ign_mode = 1, linsol = LINSOL_INVERSE;
  double ign_value = 400.0, jac_clip = 10.0;
  int gang = 1, pool_enabled = 1, pool_n = 0, pool_warps = 12, gang_groups = 1, gang_sleep_ns = 0, gang_max_polls = 0, gang_period = 1;
  bool have_ic = false, fresh = true, roberts = false, use_static_shape = true, atol_is_vec = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  DevBuf<double> T0, p0, Y0, rho0, state, save, traj, tau, tcur, scratch, tab_t, tab_T, tab_p, atol_vec, tmp_in, tmp_out, tmp_w, tmp_qf, tmp_qr, tmp_t;
  DevBuf<long long> stats;
  DevBuf<double> stats_ex;  // qlast, qcur, hinused, hlast, hcur, tcur per reactor
  DevBuf<int> status, order;
  DevBuf<unsigned long long> counter;
  DevBuf<long long> timeline; int timeline_on = 0;
  DevBuf<double> trace; int trace_cap = 0; int64_t trace_reactor = -1;
  // time slicing (see BatchArgs): ring of reactor ids, per-reactor scratch
  DevBuf<int> ring, yielded; DevBuf<unsigned long long> ring_seq, qctl;
  int slice_steps = 64, ring_cap = 0; bool sliced = false;
  const double *pT0 = nullptr, *pp0 = nullptr, *pY0 = nullptr;  // owned or borrowed device pointers
  int nt = 0;
  bool have_order = false;
  // launch configuration
  int wpb = 0, grid = 0;
  size_t smem = 0;
  double last_ms = 0.0;
  int64_t launches = 0, aux_launches = 0;  // solver kernels / k_init + k_ring_init
  ~ct_batch() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
};

namespace {

int configure_launch(ct_batch* b) {
  DeviceInfo di;
  if (int r = device_info(di)) return r;
  const MechDesc& d = b->mech ? b->mech->desc : MechDesc{};
  const WarpLayout L = make_layout(b->N, d.R);
  const size_t slab = (size_t)L.doubles * sizeof(double);
  const size_t avail = (size_t)di.max_smem_optin;
  const size_t fixed = CT_HDR_BYTES + (size_t)d.blob_bytes;
  if (fixed + slab > avail) return fail(CT_ERR_UNSUPPORTED, "mechanism too large for one SM's shared memory");
  int w = (int)((avail - fixed) / slab);
  w = std::min(w, 16);
  b->wpb = w;
  b->smem = fixed + (size_t)w * slab;
  b->pool_n = 0;
  // pool mode: Newton matrix out of the per-warp slab (shared work slabs + inverse in global scratch) when the classic
  // layout leaves fewer than 10 warps per SM; needs the explicit-inverse solver
  if (b->pool_enabled && b->mech && b->linsol == LINSOL_INVERSE && w < 10) {
    const size_t slabW = (size_t)make_layout(b->N, d.R, false).doubles * sizeof(double);
    const size_t slabM = pool_slab_bytes(b->N);
    const int pw = std::max(8, std::min(16, b->pool_warps));
    if (fixed + 16 + (size_t)pw * slabW + 3 * slabM <= avail) {
      const int pn = (int)std::min<size_t>(16, (avail - fixed - 16 - (size_t)pw * slabW) / slabM);
      b->wpb = pw; b->pool_n = pn; w = pw;
      b->smem = ((fixed + (size_t)pw * slabW + 15) & ~(size_t)15) + (size_t)pn * slabM;
    }
  }
  // persistent: one CTA per SM, never more warps than reactors
  int64_t need = (b->n + w - 1) / w;
  b->grid = (int)std::max<int64_t>(1, std::min<int64_t>(di.sms, need));
  // time slicing needs the reactor-indexed scratch (49 KB per GRI-3.0 reactor): only while that stays below 4 GiB and
  // the batch is more than one wave of warps (otherwise nothing ever waits for a warp)
  b->sliced = b->slice_steps > 0 && b->pool_n > 0 && b->trace_cap == 0 && b->n > (int64_t)b->grid * w &&
              (double)b->n * (double)warp_scratch_doubles(b->N, d.R) * 8.0 <= 4.0 * 1024 * 1024 * 1024;
  return 0;
}

BatchArgs make_args(ct_batch* b) {
  BatchArgs a{};
  a.rt = b->roberts ? RT_ROBERTS : b->rt;
  a.N = b->N; a.jac_mode = b->jac_mode; a.max_order = b->max_order; a.ign_mode = b->ign_mode; a.fresh = b->fresh ? 1 : 0;
  a.n_out = 0; a.n = b->n; a.mxstep = b->mxstep; a.linsol = b->linsol; a.jac_clip = b->jac_clip; a.gang = b->gang; a.pool_n = b->pool_n; a.gang_groups = b->gang_groups; a.gang_sleep_ns = b->gang_sleep_ns; a.gang_max_polls = b->gang_max_polls; a.gang_period = b->gang_period;
  a.T0 = b->pT0; a.p0 = b->pp0; a.rho0 = b->rho0.p;
  a.nt = b->nt; a.tab_t = b->tab_t.p; a.tab_T = b->tab_T.p; a.tab_p = b->tab_p.p;
  a.rtol = b->rtol; a.atol = b->atol; a.atol_vec = (b->atol_is_vec || b->roberts) && b->atol_vec.n ? b->atol_vec.p : nullptr;
  a.t_target = b->t_stop; a.t_stop = b->t_stop; a.ign_value = b->ign_value;
  a.state = b->state.p; a.save = nullptr; a.save_stride = CT_SAVE_DOUBLES; a.traj = nullptr;
  a.tau = b->tau.p; a.tcur = b->tcur.p; a.stats = b->stats.p; a.stats_ex = b->stats_ex.p; a.status = b->status.p; a.counter = b->counter.p;
  a.scratch = b->scratch.p; a.order = b->have_order ? b->order.p : nullptr;
  a.timeline = b->timeline_on ? b->timeline.p : nullptr;
  a.opts = b->opts; a.t_start = b->t_start;
  if (b->trace_cap > 0) { a.trace = b->trace.p; a.trace_cap = b->trace_cap; a.trace_reactor = b->trace_reactor; }
  if (b->sliced) {
    a.slice_steps = b->slice_steps; a.ring_cap = b->ring_cap; a.ring = b->ring.p; a.ring_seq = b->ring_seq.p; a.qctl = b->qctl.p;
    a.yielded = b->yielded.p; a.save = b->save.p;
  }
  return a;
}

template <class Kern>
int set_smem(Kern k, size_t bytes) {
  CT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

int ensure_runtime_buffers(ct_batch* b) {
  if (int r = configure_launch(b)) return r;
  const int R = b->mech ? b->mech->desc.R : 0;
  const size_t scr = warp_scratch_doubles(b->N, R);
  CT_CUDA(b->scratch.alloc(std::max((size_t)b->grid * b->wpb * scr, b->sliced ? (size_t)b->n * scr : (size_t)0)));
  if (b->sliced) {
    b->ring_cap = (int)std::min<int64_t>(2 * b->n, 2147483647LL);
    CT_CUDA(b->ring.alloc(b->ring_cap)); CT_CUDA(b->ring_seq.alloc(b->ring_cap)); CT_CUDA(b->qctl.alloc(4)); CT_CUDA(b->yielded.alloc(b->n));
    CT_CUDA(b->save.alloc((size_t)b->n * (CT_SAVE_DOUBLES)));
  }
  CT_CUDA(b->tau.alloc(b->n)); CT_CUDA(b->tcur.alloc(b->n)); CT_CUDA(b->stats.alloc((size_t)b->n * 8)); CT_CUDA(b->stats_ex.alloc((size_t)b->n * 6));
  CT_CUDA(b->status.alloc(b->n)); CT_CUDA(b->counter.alloc(1));
  if (!b->ev0) { CT_CUDA(cudaEventCreate(&b->ev0)); CT_CUDA(cudaEventCreate(&b->ev1)); }
  return 0;
}

int launch_integrate(ct_batch* b, BatchArgs& a) {
  const unsigned char* blob = nullptr;
  MechDesc d{};
  if (b->mech) { if (int r = b->mech->on_device(&blob, nullptr)) return r; d = b->mech->desc; }
  CT_CUDA(cudaMemsetAsync(b->counter.p, 0, sizeof(unsigned long long), b->stream));
  if (a.ring) {
    k_ring_init<<<(a.ring_cap + 255) / 256, 256, 0, b->stream>>>((int)b->n, a.ring_cap, a.order, a.ring, a.ring_seq, a.qctl, a.yielded);
    CT_CUDA(cudaGetLastError());
    b->aux_launches += 1;
  }
  CT_CUDA(cudaEventRecord(b->ev0, b->stream));
  // specialised kernels (compile-time mechanism shape) for the 12-warp pool configuration of the GRI mechanisms, the
  // run-time shape otherwise
  const int shape = (b->mech && b->use_static_shape) ? b->mech->shape : 0;
#define CT_LAUNCH(MAXT, SH)                                                            \
  do {                                                                                 \
    if (int r = set_smem(k_integrate<MAXT, SH>, b->smem)) return r;                    \
    k_integrate<MAXT, SH><<<b->grid, b->wpb * 32, b->smem, b->stream>>>(d, blob, a);   \
  } while (0)
  if (b->wpb <= 8) CT_LAUNCH(256, ShapeRt);
  else if (b->wpb <= 12) {
    if (shape == 1) CT_LAUNCH(384, ShapeGri30);
    else if (shape == 2) CT_LAUNCH(384, ShapeGri211);
    else if (shape == 3) CT_LAUNCH(384, ShapeGri12);
    else CT_LAUNCH(384, ShapeRt);
  } else CT_LAUNCH(512, ShapeRt);
#undef CT_LAUNCH
  CT_CUDA(cudaGetLastError());
  CT_CUDA(cudaEventRecord(b->ev1, b->stream));
  CT_CUDA(cudaStreamSynchronize(b->stream));
  float ms = 0.f;
  CT_CUDA(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  b->last_ms = ms;
  b->launches += 1;
  return 0;
}

int summarize_status(ct_batch* b, int* out) {
  std::vector<int> st(b->n);
  CT_CUDA(cudaMemcpy(st.data(), b->status.p, b->n * sizeof(int), cudaMemcpyDeviceToHost));
  int worst = 0;
  bool all_tstop = true, any_live = false;
  for (int s : st) {
    if (s < worst) worst = s;
    if (s != CT_TSTOP_RETURN) all_tstop = false;
    any_live = true;
  }
  *out = worst < 0 ? worst : (all_tstop && any_live ? CT_TSTOP_RETURN : CT_SUCCESS);
  return 0;
}

}  // namespace

struct ct_input {
  InputParameters P;
  std::vector<std::string> warnings;
  std::vector<InputCaseSet>& sets(int type) { return type == 0 ? P.const_P : type == 1 ? P.const_V : type == 2 ? P.const_TP : P.const_TV; }
  const InputCaseSet* set(int type, int i) {
    if (type < 0 || type > 3) return nullptr;
    auto& v = sets(type);
    return (i >= 0 && i < (int)v.size()) ? &v[i] : nullptr;
  }
  const InputCase* cas(int type, int i, int c) {
    const InputCaseSet* s = set(type, i);
    return (s && c >= 0 && c < (int)s->cases.size()) ? &s->cases[c] : nullptr;
  }
};

extern "C" {

const char* ct_last_error(void) { return g_err.c_str(); }

// ---------------------------------------------------------------- results (Results::HDF5Writer layout)
struct ct_h5 { H5File f; };
ct_h5* ct_h5_create(void) { return new ct_h5; }
void ct_h5_destroy(ct_h5* h) { delete h; }
int ct_h5_add_dataset(ct_h5* h, const char* path, int rank, const int64_t* dims, const double* data, const char* name,
                      const char* notation, const char* unit) {
  if (!h || !path || rank < 0 || rank > 4 || (rank > 0 && !dims) || !data) return fail(CT_ERR_INVALID, "bad dataset arguments");
  try {
    std::vector<uint64_t> d(rank);
    for (int i = 0; i < rank; ++i) { if (dims[i] < 0) return fail(CT_ERR_INVALID, "negative dimension"); d[i] = (uint64_t)dims[i]; }
    std::vector<std::pair<std::string, std::string>> attrs;
    // Results::Meta::Register(group, set, name, notation, unit, ptr) (include/results/observer.hpp:59-72)
    if (name) attrs.emplace_back("Name", name);
    if (notation) attrs.emplace_back("Notation", notation);
    if (unit) attrs.emplace_back("Unit", unit);
    h->f.create_dataset(path, d, data, attrs);
  } catch (const std::exception& e) { return fail(CT_ERR_INVALID, e.what()); }
  return 0;
}
int ct_h5_set_layout(ct_h5* h, int64_t chunk_elems, int deflate_level, int vlen_strings) {
  if (!h || chunk_elems < 0) return fail(CT_ERR_INVALID, "bad argument");
  try { h->f.set_layout((uint64_t)chunk_elems, deflate_level, vlen_strings != 0); } catch (const std::exception& e) { return fail(CT_ERR_INVALID, e.what()); }
  return 0;
}
int ct_h5_write(ct_h5* h, const char* filename) {
  if (!h || !filename) return fail(CT_ERR_INVALID, "null argument");
  try { h->f.write(filename); } catch (const std::exception& e) { return fail(CT_ERR_FILESYSTEM, e.what()); }
  return 0;
}

// ---------------------------------------------------------------- YAML input (Input::Parser)
ct_input* ct_input_parse(const char* config_file, int* status) {
  auto set = [&](int s) { if (status) *status = s; };
  if (!config_file) { set(fail(CT_ERR_INVALID, "null config path")); return nullptr; }
  try {
    std::unique_ptr<ct_input> in(new ct_input);
    in->P = parse_input(config_file);
    // Parser::Validate (src/input_parser.cpp:310-351): mechanism file exists, then every case against its thermo
    for (int t = 0; t < 4; ++t)
      for (auto& cs : in->sets(t)) {
        struct stat sb;
        if (stat(cs.mechanism_path.c_str(), &sb) != 0) throw InputError(InputError::Filesystem, "Input error: " + cs.mechanism_path + ": No such file or directory");
        const bool is_ctm = cs.mechanism_path.size() > 4 && cs.mechanism_path.compare(cs.mechanism_path.size() - 4, 4, ".ctm") == 0;
        MechanismDef def = is_ctm ? read_ctm(cs.mechanism_path) : parse_cti(cs.mechanism_path, nullptr);
        CompiledMech C = compile_mechanism(def, "");
        double Tmin = 0.0, Tmax = 1e300;  // ThermoPhase::minTemp / maxTemp
        for (auto& sp : def.species) { Tmin = std::max(Tmin, sp.Tlo); Tmax = std::min(Tmax, sp.Thi); }
        for (auto& c : cs.cases) validate_case(c, C, Tmin, Tmax, &in->warnings);
      }
    set(0);
    return in.release();
  } catch (const InputError& e) {
    static const int codes[] = {CT_ERR_INVALID, CT_ERR_YAML_BAD_FILE, CT_ERR_YAML_INVALID_NODE, CT_ERR_YAML_BAD_CONVERSION, CT_ERR_FILESYSTEM, CT_ERR_INPUT, CT_ERR_INPUT_UNIT};
    set(fail(codes[e.kind], e.what()));
  } catch (const std::exception& e) {
    set(fail(CT_ERR_MECH, e.what()));
  }
  return nullptr;
}
void ct_input_destroy(ct_input* in) { delete in; }
int ct_input_n_case_sets(ct_input* in, int type) { return (in && type >= 0 && type <= 3) ? (int)in->sets(type).size() : fail(CT_ERR_INVALID, "bad argument"); }
int ct_input_n_cases(ct_input* in, int type, int set) { auto* s = in ? in->set(type, set) : nullptr; return s ? (int)s->cases.size() : fail(CT_ERR_INVALID, "bad case set"); }
const char* ct_input_mechanism_path(ct_input* in, int type, int set) { auto* s = in ? in->set(type, set) : nullptr; return s ? s->mechanism_path.c_str() : nullptr; }
const char* ct_input_mechanism_description(ct_input* in, int type, int set) { auto* s = in ? in->set(type, set) : nullptr; return (s && s->has_mechanism_description) ? s->mechanism_description.c_str() : nullptr; }
const char* ct_input_set_description(ct_input* in, int type, int set) { auto* s = in ? in->set(type, set) : nullptr; return (s && s->has_description) ? s->description.c_str() : nullptr; }
const char* ct_input_case_description(ct_input* in, int type, int set, int c) { auto* k = in ? in->cas(type, set, c) : nullptr; return (k && k->has_description) ? k->description.c_str() : nullptr; }
int ct_input_case(ct_input* in, int type, int set, int c, double* pressure, double* temperature, double* time, int* comp_type, int* n_pairs) {
  auto* k = in ? in->cas(type, set, c) : nullptr;
  if (!k) return fail(CT_ERR_INVALID, "bad case index");
  if (pressure) *pressure = k->pressure;
  if (temperature) *temperature = k->temperature;
  if (time) *time = k->time;
  if (comp_type) *comp_type = (int)k->composition_type;
  if (n_pairs) *n_pairs = (int)k->composition.size();
  return 0;
}
const char* ct_input_case_species(ct_input* in, int type, int set, int c, int i, double* value) {
  auto* k = in ? in->cas(type, set, c) : nullptr;
  if (!k || i < 0 || i >= (int)k->composition.size()) return nullptr;
  if (value) *value = k->composition[i].second;
  return k->composition[i].first.c_str();
}
int ct_input_case_mass_fractions(ct_input* in, int type, int set, int c, const ct_mech* m, double* Y) {
  auto* k = in ? in->cas(type, set, c) : nullptr;
  if (!k || !m || !Y) return fail(CT_ERR_INVALID, "bad argument");
  try {
    std::vector<double> y = case_mass_fractions(*k, m->C);
    std::copy(y.begin(), y.end(), Y);
  } catch (const std::exception& e) { return fail(CT_ERR_INPUT, e.what()); }
  return 0;
}
int ct_input_n_warnings(ct_input* in) { return in ? (int)in->warnings.size() : 0; }
const char* ct_input_warning(ct_input* in, int i) { return (in && i >= 0 && i < (int)in->warnings.size()) ? in->warnings[i].c_str() : nullptr; }

const char* ct_version(void) { return "chemtools_b200 0.1 (sm_100a)"; }

ct_mech* ct_mech_load(const char* path, const char* phase_id, int* status) { return ct_mech_load_ex(path, phase_id, nullptr, status); }

ct_mech* ct_mech_load_ex(const char* path, const char* phase_id, const char* constants, int* status) {
  auto set = [&](int s) { if (status) *status = s; };
  if (!path) { set(fail(CT_ERR_INVALID, "null mechanism path")); return nullptr; }
  try {
    std::unique_ptr<ct_mech> m(new ct_mech);
    const std::string p(path);
    const bool is_ctm = p.size() > 4 && p.compare(p.size() - 4, 4, ".ctm") == 0;
    m->def = is_ctm ? read_ctm(p) : parse_cti(p, phase_id);
    m->C = compile_mechanism(m->def, constants ? constants : "");
    m->desc = make_desc(m->C);
    m->shape = static_shape_of(m->desc);
    if (m->C.K + 1 > CT_NP) { set(fail(CT_ERR_UNSUPPORTED, "more than 63 species: the warp-per-reactor kernels hold 2 state components per lane")); return nullptr; }
    if (m->C.n_fall + m->C.n_3b > CT_NP) { set(fail(CT_ERR_UNSUPPORTED, "more than 64 third-body/falloff reactions")); return nullptr; }
    set(0);
    return m.release();
  } catch (const std::exception& e) {
    set(fail(CT_ERR_MECH, e.what()));
    return nullptr;
  }
}
void ct_mech_destroy(ct_mech* m) { delete m; }
int ct_mech_save_ctm(const ct_mech* m, const char* path) {
  if (!m || !path) return fail(CT_ERR_INVALID, "null argument");
  try { write_ctm(m->def, path); } catch (const std::exception& e) { return fail(CT_ERR_MECH, e.what()); }
  return 0;
}
int ct_mech_n_species(const ct_mech* m) { return m ? m->C.K : fail(CT_ERR_INVALID, "null mech"); }
int ct_mech_n_reactions(const ct_mech* m) { return m ? m->C.R : fail(CT_ERR_INVALID, "null mech"); }
int ct_mech_n_elements(const ct_mech* m) { return m ? (int)m->C.elements.size() : fail(CT_ERR_INVALID, "null mech"); }
int ct_mech_species_index(const ct_mech* m, const char* name) {
  if (!m || !name) return -1;
  for (int k = 0; k < m->C.K; ++k) if (m->C.species_names[k] == name) return k;
  return -1;
}
const char* ct_mech_species_name(const ct_mech* m, int k) { return (m && k >= 0 && k < m->C.K) ? m->C.species_names[k].c_str() : nullptr; }
const char* ct_mech_reaction_equation(const ct_mech* m, int i) { return (m && i >= 0 && i < m->C.R) ? m->C.equations[i].c_str() : nullptr; }
const char* ct_mech_element_name(const ct_mech* m, int e) { return (m && e >= 0 && e < (int)m->C.elements.size()) ? m->C.elements[e].c_str() : nullptr; }
double ct_mech_n_atoms(const ct_mech* m, int k, int e) {
  if (!m || k < 0 || k >= m->C.K || e < 0 || e >= (int)m->C.elements.size()) return 0.0;
  return m->C.natoms[(size_t)k * m->C.elements.size() + e];
}
int ct_mech_molecular_weights(const ct_mech* m, double* W) {
  if (!m || !W) return fail(CT_ERR_INVALID, "null argument");
  std::copy(m->C.W.begin(), m->C.W.end(), W);
  return 0;
}
double ct_mech_gas_constant(const ct_mech* m) { return m ? m->C.gas_constant : 0.0; }
int ct_mech_compiled_arrays(const ct_mech* m, int* perm, double* A, double* b, double* EaR) {
  if (!m) return fail(CT_ERR_INVALID, "null mech");
  const auto& C = m->C;
  const unsigned char* base = C.blob.data();
  if (perm) std::copy(C.perm.begin(), C.perm.end(), perm);
  if (A) std::memcpy(A, base + C.off.A, C.R * sizeof(double));
  if (b) std::memcpy(b, base + C.off.b, C.R * sizeof(double));
  if (EaR) std::memcpy(EaR, base + C.off.E, C.R * sizeof(double));
  return 0;
}
int ct_mech_mole_to_mass(const ct_mech* m, int64_t n, const double* X, double* Y) {
  if (!m || !X || !Y || n < 0) return fail(CT_ERR_INVALID, "bad argument");
  const int K = m->C.K;
  for (int64_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += X[i * K + k] * m->C.W[k];
    if (!(s > 0.0)) return fail(CT_ERR_INVALID, "composition sums to zero");
    for (int k = 0; k < K; ++k) Y[i * K + k] = X[i * K + k] * m->C.W[k] / s;
  }
  return 0;
}
int ct_mech_flop_model(const ct_mech* m, int reactor_type, double out[4]) {
  if (!m || !out) return fail(CT_ERR_INVALID, "null argument");
  const int N = rt_has_energy(reactor_type) ? m->C.K + 1 : m->C.K;
  out[0] = m->C.flops_rhs; out[1] = m->C.flops_jac;
  out[2] = 2.0 / 3.0 * (double)N * N * N; out[3] = 2.0 * (double)N * N;
  return 0;
}

int ct_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
int ct_set_device(int ordinal) { if (int r = require_device()) return r; CT_CUDA(cudaSetDevice(ordinal)); return 0; }

ct_batch* ct_batch_create(ct_mech* m, int rt, int64_t n, int* status) {
  auto set = [&](int s) { if (status) *status = s; };
  if (!m || n < 0 || rt < 0 || rt > CT_TABLE_V) { set(fail(CT_ERR_INVALID, "bad batch arguments")); return nullptr; }
  if (int r = require_device()) { set(r); return nullptr; }
  ct_batch* b = new ct_batch;
  b->mech = m; b->rt = rt; b->n = n; b->K = m->C.K; b->N = rt_has_energy(rt) ? m->C.K + 1 : m->C.K;
  set(0);
  return b;
}
void ct_batch_destroy(ct_batch* b) { delete b; }
int ct_batch_n_state(const ct_batch* b) { return b ? b->N : fail(CT_ERR_INVALID, "null batch"); }
int ct_batch_set_stream(ct_batch* b, void* s) { if (!b) return fail(CT_ERR_INVALID, "null batch"); b->stream = (cudaStream_t)s; return 0; }

static int finish_ic(ct_batch* b) {
  const unsigned char* blob = nullptr;
  if (int r = b->mech->on_device(&blob, nullptr)) return r;
  CT_CUDA(b->rho0.alloc(b->n));
  CT_CUDA(b->state.alloc((size_t)b->n * b->N));
  if (b->n > 0) {
    const int threads = 128;
    const int blocks = (int)std::min<int64_t>((b->n + threads - 1) / threads, 4096);
    k_init<<<blocks, threads, 0, b->stream>>>(b->mech->desc, blob, b->rt, b->n, b->pT0, b->pp0, b->pY0, b->state.p, b->rho0.p, b->N);
    CT_CUDA(cudaGetLastError());
    b->aux_launches += 1;
  }
  b->have_ic = true; b->fresh = true; b->t_start = 0.0;
  return 0;
}
int ct_batch_set_ic(ct_batch* b, const double* T, const double* p, const double* Y) {
  if (!b || !T || !p || !Y) return fail(CT_ERR_INVALID, "null argument");
  CT_CUDA(b->T0.alloc(b->n)); CT_CUDA(b->p0.alloc(b->n)); CT_CUDA(b->Y0.alloc((size_t)b->n * b->K));
  CT_CUDA(cudaMemcpyAsync(b->T0.p, T, b->n * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  CT_CUDA(cudaMemcpyAsync(b->p0.p, p, b->n * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  CT_CUDA(cudaMemcpyAsync(b->Y0.p, Y, (size_t)b->n * b->K * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  b->pT0 = b->T0.p; b->pp0 = b->p0.p; b->pY0 = b->Y0.p;
  return finish_ic(b);
}
int ct_batch_set_ic_device(ct_batch* b, const double* dT, const double* dp, const double* dY) {
  if (!b || !dT || !dp || !dY) return fail(CT_ERR_INVALID, "null argument");
  b->pT0 = dT; b->pp0 = dp; b->pY0 = dY;
  return finish_ic(b);
}
int ct_batch_set_tolerances(ct_batch* b, double rtol, double atol) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  // same validation as IStrategy::Initialise (interface.cpp:300-317): negative tolerances are illegal
  if (!(rtol >= 0.0) || !(atol >= 0.0) || (rtol == 0.0 && atol == 0.0)) return fail(CT_ILL_INPUT, "tolerances must be non-negative and not both zero");
  b->rtol = rtol; b->atol = atol; b->atol_is_vec = false;
  return 0;
}
// Client::SetTolerances(rtol, vector atol) (client.hpp:122-134, types.hpp:114-138) -> CVodeSVtolerances: one absolute
// tolerance per state component (N entries: [T, Y_0..] or [Y_0..]), shared by all reactors of the batch
int ct_batch_set_tolerances_vec(ct_batch* b, double rtol, const double* atol) {
  if (!b || !atol) return fail(CT_ERR_INVALID, "null argument");
  if (!(rtol >= 0.0)) return fail(CT_ILL_INPUT, "relative tolerance must be non-negative");
  bool any = rtol > 0.0;
  for (int i = 0; i < b->N; ++i) {
    if (!(atol[i] >= 0.0)) return fail(CT_ILL_INPUT, "absolute tolerances must be non-negative");
    any = any || atol[i] > 0.0;
  }
  if (!any) return fail(CT_ILL_INPUT, "tolerances must not all be zero");
  CT_CUDA(b->atol_vec.alloc(b->N));
  CT_CUDA(cudaMemcpyAsync(b->atol_vec.p, atol, b->N * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  CT_CUDA(cudaStreamSynchronize(b->stream));
  b->rtol = rtol; b->atol = 0.0; b->atol_is_vec = true;
  return 0;
}
int ct_batch_set_stop_time(ct_batch* b, double t_end) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (!(t_end > 0.0)) return fail(CT_ILL_INPUT, "stop time must be positive");
  b->t_stop = t_end;
  return 0;
}
// IStrategy::SetMaxNumSteps (interface.cpp:100-106): negative values are refused; 0 selects CVODE's default of 500
int ct_batch_set_max_steps(ct_batch* b, int64_t mx) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (mx < 0) return fail(CT_ILL_INPUT, "SetMaxNumSteps: maximum number of steps must be non-negative");
  b->mxstep = mx > 0 ? std::min<int64_t>(mx, 2147483647LL) : 500;  // the device counts steps in 32 bits
  return 0;
}
int ct_batch_set_max_order(ct_batch* b, int q) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (q < 1 || q > 5) return fail(CT_ILL_INPUT, "BDF order must be in 1..5");
  b->max_order = q;
  return 0;
}
// IStrategy::SetInitStepSize / SetMinStepSize / SetMaxStepSize (interface.cpp:61-97): negative values are invalid,
// 0 selects CVODE's default (estimated first step, hmin = 0, hmax = infinity)
int ct_batch_set_init_step(ct_batch* b, double h) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (!(h >= 0.0)) return fail(CT_ILL_INPUT, "SetInitStepSize: initial step size must be non-negative");
  b->opts.h_init = h;
  return 0;
}
int ct_batch_set_min_step(ct_batch* b, double h) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (!(h >= 0.0)) return fail(CT_ILL_INPUT, "SetMinStepSize: minimum step size must be non-negative");
  if (h > 0.0 && b->opts.h_max_inv > 0.0 && h * b->opts.h_max_inv > 1.0) return fail(CT_ILL_INPUT, "SetMinStepSize: minimum step exceeds maximum step");
  b->opts.h_min = h;
  return 0;
}
int ct_batch_set_max_step(ct_batch* b, double h) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (!(h >= 0.0)) return fail(CT_ILL_INPUT, "SetMaxStepSize: maximum step size must be non-negative");
  if (h > 0.0 && h < b->opts.h_min) return fail(CT_ILL_INPUT, "SetMaxStepSize: maximum step below minimum step");
  b->opts.h_max_inv = h > 0.0 ? 1.0 / h : 0.0;
  return 0;
}
// IStrategy::SetMaxErrorTestFailures / SetMaxNonlinearIterations / SetMaxConvergenceFailures /
// SetNonlinConvergenceCoefficient (interface.cpp:166-219): must be positive; defaults 7 / 3 / 10 / 0.1 (types.hpp:97-111)
int ct_batch_set_max_err_test_fails(ct_batch* b, int n) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (n <= 0) return fail(CT_ILL_INPUT, "SetMaxErrorTestFailures: must be positive");
  b->opts.max_nef = n;
  return 0;
}
int ct_batch_set_max_nonlin_iters(ct_batch* b, int n) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (n <= 0) return fail(CT_ILL_INPUT, "SetMaxNonlinearIterations: must be positive");
  b->opts.max_cor = n;
  return 0;
}
int ct_batch_set_max_conv_fails(ct_batch* b, int n) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (n <= 0) return fail(CT_ILL_INPUT, "SetMaxConvergenceFailures: must be positive");
  b->opts.max_ncf = n;
  return 0;
}
int ct_batch_set_nonlin_conv_coef(ct_batch* b, double c) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (!(c > 0.0)) return fail(CT_ILL_INPUT, "SetNonlinConvergenceCoefficient: must be positive");
  b->opts.nls_coef = c;
  return 0;
}
// IStrategy::SetMaxWarnMessages (interface.cpp:125-137): CVODE's "t + h = t" warnings are not printed by the device
// stepper, the count is validated and otherwise ignored.  IStrategy::SetStabilityLimitDetection (interface.cpp:139-164):
// the BDF stability-limit detection algorithm is not implemented; switching it on is refused instead of ignored.
int ct_batch_set_max_warn_messages(ct_batch* b, int n) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (n < 0) return fail(CT_ILL_INPUT, "SetMaxWarnMessages: must be non-negative");
  return 0;
}
int ct_batch_set_stability_limit_detection(ct_batch* b, int on) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (on) return fail(CT_ERR_UNSUPPORTED, "stability limit detection (CVodeSetStabLimDet) is not implemented");
  return 0;
}
// IStrategy::ReInitialise(time_init, state) (interface.cpp:429-446) -> CVodeReInit: new start time and state, all
// options, tolerances and the reactors' frozen parameters (pressure / density of the original initial conditions) kept
int ct_batch_reinit(ct_batch* b, double t0, const double* state) {
  if (!b || !state) return fail(CT_ERR_INVALID, "null argument");
  if (!b->have_ic) return fail(CT_ILL_INPUT, "ReInitialise before the initial conditions were set");
  if (!(t0 < b->t_stop)) return fail(CT_ILL_INPUT, "ReInitialise: start time must precede the stop time");
  if (b->n > 0) CT_CUDA(cudaMemcpyAsync(b->state.p, state, (size_t)b->n * b->N * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  CT_CUDA(cudaStreamSynchronize(b->stream));
  b->t_start = t0; b->fresh = true;
  return 0;
}
int ct_batch_set_jacobian(ct_batch* b, int mode) {
  if (!b || (mode != CT_JAC_DQ && mode != CT_JAC_ANALYTIC)) return fail(CT_ERR_INVALID, "bad Jacobian mode");
  b->jac_mode = mode;
  return 0;
}
int ct_batch_set_linear_solver(ct_batch* b, int mode) {
  if (!b || (mode != LINSOL_LU && mode != LINSOL_INVERSE)) return fail(CT_ERR_INVALID, "bad linear solver mode");
  if (!b->fresh) return fail(CT_ILL_INPUT, "the linear solver can only change on a fresh batch (set the initial conditions again)");
  b->linsol = mode;
  return 0;
}
// the scratch / save-area indexing of a batch that has been integrated part of the way (Integrate(t < t_stop)) depends on
// these options: they may only change on a fresh batch (after set_ic / reinit)
#define CT_REQUIRE_FRESH(b) do { if (!(b)->fresh) return fail(CT_ILL_INPUT, "scheduling layout options can only change on a fresh batch (set the initial conditions again)"); } while (0)
int ct_batch_set_matrix_pool(ct_batch* b, int on, int warps_per_sm) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  CT_REQUIRE_FRESH(b);
  b->pool_enabled = on ? 1 : 0;
  if (warps_per_sm > 0) b->pool_warps = warps_per_sm;
  return 0;
}
int ct_batch_set_static_shape(ct_batch* b, int on) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  b->use_static_shape = on != 0;
  return 0;
}
int ct_mech_static_shape(const ct_mech* m) { return m ? m->shape : fail(CT_ERR_INVALID, "null mech"); }
int ct_batch_set_phase_alignment(ct_batch* b, int on) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  b->gang = on ? 1 : 0;
  if (on > 1) b->gang_groups = 2;  // two alignment groups (even / odd warps)
  else b->gang_groups = 1;
  return 0;
}
int ct_batch_set_alignment_poll_ns(ct_batch* b, int ns) {
  if (!b || ns < 0) return fail(CT_ERR_INVALID, "bad argument");
  b->gang_sleep_ns = ns;
  return 0;
}
int ct_batch_set_alignment_period(ct_batch* b, int period) {
  if (!b || period < 1) return fail(CT_ERR_INVALID, "bad argument");
  b->gang_period = period;
  return 0;
}
int ct_batch_set_alignment_max_polls(ct_batch* b, int polls) {
  if (!b || polls < 0) return fail(CT_ERR_INVALID, "bad argument");
  b->gang_max_polls = polls;
  return 0;
}
int ct_batch_set_time_slicing(ct_batch* b, int steps) {
  if (!b || steps < 0) return fail(CT_ERR_INVALID, "bad argument");
  CT_REQUIRE_FRESH(b);
  b->slice_steps = steps;
  return 0;
}
int ct_batch_set_trace(ct_batch* b, int64_t reactor, int capacity) {
  if (!b || capacity < 0 || (capacity > 0 && (reactor < 0 || reactor >= b->n))) return fail(CT_ERR_INVALID, "bad trace arguments");
  b->trace_cap = capacity; b->trace_reactor = reactor;
  if (capacity > 0) {
    CT_CUDA(b->trace.alloc((size_t)capacity * CT_TRACE_W + 1));
    CT_CUDA(cudaMemset(b->trace.p, 0, ((size_t)capacity * CT_TRACE_W + 1) * sizeof(double)));
  }
  return 0;
}
int ct_batch_get_trace(ct_batch* b, double* out, int* count) {
  if (!b || !out || !count) return fail(CT_ERR_INVALID, "null argument");
  if (b->trace_cap <= 0) return fail(CT_ERR_INVALID, "step trace is off");
  CT_CUDA(cudaStreamSynchronize(b->stream));
  std::vector<double> h((size_t)b->trace_cap * CT_TRACE_W + 1);
  CT_CUDA(cudaMemcpy(h.data(), b->trace.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
  *count = (int)h.back();
  std::memcpy(out, h.data(), (size_t)b->trace_cap * CT_TRACE_W * sizeof(double));
  return 0;
}
int ct_batch_set_timeline(ct_batch* b, int on) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (on) { CT_CUDA(b->timeline.alloc((size_t)b->n * 3)); CT_CUDA(cudaMemset(b->timeline.p, 0, (size_t)b->n * 3 * sizeof(long long))); }
  b->timeline_on = on ? 1 : 0;
  return 0;
}
int ct_batch_get_timeline(ct_batch* b, int64_t* out) {
  if (!b || !out) return fail(CT_ERR_INVALID, "null argument");
  if (!b->timeline_on || !b->timeline.p) return fail(CT_ERR_INVALID, "timeline recording is off");
  CT_CUDA(cudaStreamSynchronize(b->stream));
  CT_CUDA(cudaMemcpy(out, b->timeline.p, (size_t)b->n * 3 * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}
int ct_batch_set_jacobian_clip(ct_batch* b, double thr) {
  if (!b || !(thr >= 0.0)) return fail(CT_ERR_INVALID, "bad clip threshold");
  b->jac_clip = thr;
  return 0;
}
int ct_batch_set_ignition(ct_batch* b, int mode, double value) {
  if (!b || (mode != 0 && mode != 1)) return fail(CT_ERR_INVALID, "bad ignition mode");
  b->ign_mode = mode; b->ign_value = value;
  return 0;
}
static int set_table(ct_batch* b, int nt, const double* t, const double* T, const double* p);
int ct_batch_set_tp_table(ct_batch* b, int nt, const double* t, const double* T, const double* p) {
  if (!b || nt < 2 || !t || !T || !p) return fail(CT_ERR_INVALID, "bad table arguments");
  if (b->rt != CT_TABLE_TP) return fail(CT_ERR_INVALID, "T-p table on a reactor type other than CT_TABLE_TP");
  return set_table(b, nt, t, T, p);
}
// prescribed volume and its time derivative (the reference's plan, README.md:18-20): V[n x nt] > 0 in any unit (only
// V(t) / V(t[0]) enters), dVdt[n x nt] in that unit per second; both interpolated linearly on the common grid
int ct_batch_set_v_table(ct_batch* b, int nt, const double* t, const double* V, const double* dVdt) {
  if (!b || nt < 2 || !t || !V || !dVdt) return fail(CT_ERR_INVALID, "bad table arguments");
  if (b->rt != CT_TABLE_V) return fail(CT_ERR_INVALID, "volume table on a reactor type other than CT_TABLE_V");
  for (int64_t i = 0; i < b->n * (int64_t)nt; ++i) if (!(V[i] > 0.0)) return fail(CT_ILL_INPUT, "volumes must be positive");
  return set_table(b, nt, t, V, dVdt);
}
static int set_table(ct_batch* b, int nt, const double* t, const double* T, const double* p) {
  for (int i = 1; i < nt; ++i) if (!(t[i] > t[i - 1])) return fail(CT_ILL_INPUT, "table times must increase strictly");
  CT_CUDA(b->tab_t.alloc(nt)); CT_CUDA(b->tab_T.alloc((size_t)b->n * nt)); CT_CUDA(b->tab_p.alloc((size_t)b->n * nt));
  CT_CUDA(cudaMemcpyAsync(b->tab_t.p, t, nt * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  CT_CUDA(cudaMemcpyAsync(b->tab_T.p, T, (size_t)b->n * nt * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  CT_CUDA(cudaMemcpyAsync(b->tab_p.p, p, (size_t)b->n * nt * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  b->nt = nt;
  return 0;
}
int ct_batch_set_order(ct_batch* b, const int32_t* order) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (!order) { b->have_order = false; return 0; }
  // must be a permutation of 0..n-1: an out-of-range entry would index past the state arrays, a duplicate would make
  // two warps integrate the same reactor
  {
    std::vector<bool> seen((size_t)b->n, false);
    for (int64_t i = 0; i < b->n; ++i) {
      const int32_t o = order[i];
      if (o < 0 || (int64_t)o >= b->n || seen[(size_t)o]) return fail(CT_ERR_INVALID, "order must be a permutation of 0..n-1");
      seen[(size_t)o] = true;
    }
  }
  CT_CUDA(b->order.alloc(b->n));
  CT_CUDA(cudaMemcpyAsync(b->order.p, order, b->n * sizeof(int), cudaMemcpyHostToDevice, b->stream));
  CT_CUDA(cudaStreamSynchronize(b->stream));  // the caller's array may go away
  b->have_order = true;
  return 0;
}

static int precheck_integrate(ct_batch* b) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (!b->have_ic) return fail(CT_ILL_INPUT, "initial conditions not set");
  if (rt_has_table(b->rt) && b->nt < 2) return fail(CT_ILL_INPUT, "table reactor (CT_TABLE_TP / CT_TABLE_V) without its table");
  return ensure_runtime_buffers(b);
}

int ct_batch_integrate(ct_batch* b, double t_target) {
  if (int r = precheck_integrate(b)) return r;
  if (b->n == 0) return CT_SUCCESS;
  if (b->fresh && !(t_target > b->t_start)) return fail(CT_ILL_INPUT, "t_target must follow the start time");
  BatchArgs a = make_args(b);
  a.t_target = t_target;
  // every Integrate leaves the stepper of each reactor in the save area (3.5 KB per reactor): a later call continues from
  // it, whether the reactor returned CV_SUCCESS before the stop time or CV_TOO_MUCH_WORK at it (CVode called again)
  CT_CUDA(b->save.alloc((size_t)b->n * a.save_stride));
  a.save = b->save.p; a.user_save = 1;
  if (b->fresh) CT_CUDA(cudaMemsetAsync(b->status.p, 0, b->n * sizeof(int), b->stream));
  else {
    // reactors that returned CV_SUCCESS at the previous output stay live: status back to 0 happens on device
  }
  if (int r = launch_integrate(b, a)) return r;
  b->fresh = false;
  int out = 0;
  if (int r = summarize_status(b, &out)) return r;
  return out;
}

}  // extern "C" (the ladder below is internal)

namespace {

// Recovery ladder over the reactors `bad` of batch b (indices into b): they are re-run from their initial conditions under
// other Jacobian / Newton options (Jacobian clip 1; LU instead of the inverse; difference-quotient Jacobian + LU, the
// reference's own configuration) and each takes the result of the FIRST rung, in that order, that brings it to the stop
// time.  About 0.05-0.2 % of stiff CH4 ensembles crawl into CV_TOO_MUCH_WORK / CV_ERR_FAILURE: a property of the clipped
// right-hand side shared with the reference's algorithm, chaotic in which reactor it hits, so another attempt
// under different arithmetic almost always gets through.  While three times the set fits one wave of warps the three
// rungs run CONCURRENTLY (three small batches, three streams, three host threads): the pass costs its slowest rung instead
// of their sum; larger sets go rung by rung, each rung only over what the previous one left.  On return `bad` holds the
// reactors no rung could finish.
int retry_ladder(ct_batch* b, std::vector<int>& bad) {
  if (bad.empty()) return 0;
  if (b->t_start != 0.0) return fail(CT_ERR_UNSUPPORTED, "IntegrateWithRetry after ReInitialise is not supported");
  struct Alt { double clip; int linsol, jac; };
  const Alt ladder[3] = {{1.0, LINSOL_INVERSE, CT_JAC_ANALYTIC}, {b->jac_clip, LINSOL_LU, CT_JAC_ANALYTIC}, {b->jac_clip, LINSOL_LU, CT_JAC_DQ}};
  const int K = b->K, N = b->N;
  int dev = 0;
  CT_CUDA(cudaGetDevice(&dev));
  DeviceInfo di;
  if (int r = device_info(di)) return r;
  const int thr = 256;
  auto blocks = [&](int64_t work) { return (int)std::max<int64_t>(1, std::min<int64_t>((work + thr - 1) / thr, 2048)); };
  int rung = 0;
  while (rung < 3 && !bad.empty()) {
    const int m = (int)bad.size();
    const int nconc = ((int64_t)(3 - rung) * m <= (int64_t)di.sms * 12) ? 3 - rung : 1;  // rungs run at once
    // initial conditions of the set, gathered once on the device and shared by the concurrent rungs
    DevBuf<int> idx;
    DevBuf<double> gT, gp, gY, gtT, gtp;
    CT_CUDA(idx.alloc(m)); CT_CUDA(gT.alloc(m)); CT_CUDA(gp.alloc(m)); CT_CUDA(gY.alloc((size_t)m * K));
    CT_CUDA(cudaMemcpyAsync(idx.p, bad.data(), m * sizeof(int), cudaMemcpyHostToDevice, b->stream));
    k_gather_rows<double><<<blocks(m), thr, 0, b->stream>>>(b->pT0, idx.p, m, 1, gT.p);
    k_gather_rows<double><<<blocks(m), thr, 0, b->stream>>>(b->pp0, idx.p, m, 1, gp.p);
    k_gather_rows<double><<<blocks((int64_t)m * K), thr, 0, b->stream>>>(b->pY0, idx.p, m, K, gY.p);
    if (rt_has_table(b->rt)) {
      CT_CUDA(gtT.alloc((size_t)m * b->nt)); CT_CUDA(gtp.alloc((size_t)m * b->nt));
      k_gather_rows<double><<<blocks((int64_t)m * b->nt), thr, 0, b->stream>>>(b->tab_T.p, idx.p, m, b->nt, gtT.p);
      k_gather_rows<double><<<blocks((int64_t)m * b->nt), thr, 0, b->stream>>>(b->tab_p.p, idx.p, m, b->nt, gtp.p);
    }
    CT_CUDA(cudaGetLastError());
    CT_CUDA(cudaStreamSynchronize(b->stream));
    b->aux_launches += 3;
    struct Run { std::unique_ptr<ct_batch> r; cudaStream_t s = nullptr; int rc = 0; std::string err; std::vector<int> st; };
    std::vector<Run> runs(nconc);
    for (int c = 0; c < nconc; ++c) {
      int stc = 0;
      runs[c].r.reset(ct_batch_create(b->mech, b->rt, m, &stc));
      if (!runs[c].r) return stc;
      ct_batch* r = runs[c].r.get();
      if (nconc > 1) { CT_CUDA(cudaStreamCreateWithFlags(&runs[c].s, cudaStreamNonBlocking)); r->stream = runs[c].s; } else r->stream = b->stream;
      r->rtol = b->rtol; r->atol = b->atol; r->t_stop = b->t_stop; r->mxstep = b->mxstep; r->opts = b->opts;
      r->max_order = b->max_order; r->ign_mode = b->ign_mode; r->ign_value = b->ign_value; r->use_static_shape = b->use_static_shape;
      r->jac_clip = ladder[rung + c].clip; r->linsol = ladder[rung + c].linsol; r->jac_mode = ladder[rung + c].jac;
      if (b->atol_is_vec) {
        CT_CUDA(r->atol_vec.alloc(N));
        CT_CUDA(cudaMemcpy(r->atol_vec.p, b->atol_vec.p, N * sizeof(double), cudaMemcpyDeviceToDevice));
        r->atol_is_vec = true;
      }
      if (rt_has_table(b->rt)) {  // copies of the gathered tables
        CT_CUDA(r->tab_t.alloc(b->nt));
        CT_CUDA(cudaMemcpy(r->tab_t.p, b->tab_t.p, b->nt * sizeof(double), cudaMemcpyDeviceToDevice));
        CT_CUDA(r->tab_T.alloc((size_t)m * b->nt)); CT_CUDA(r->tab_p.alloc((size_t)m * b->nt));
        CT_CUDA(cudaMemcpy(r->tab_T.p, gtT.p, (size_t)m * b->nt * sizeof(double), cudaMemcpyDeviceToDevice));
        CT_CUDA(cudaMemcpy(r->tab_p.p, gtp.p, (size_t)m * b->nt * sizeof(double), cudaMemcpyDeviceToDevice));
        r->nt = b->nt;
      }
    }
    auto one = [&](int c) {
      Run& R = runs[c];
      if (cudaSetDevice(dev) != cudaSuccess) { R.rc = CT_ERR_CUDA; R.err = "cudaSetDevice"; return; }
      int rc = ct_batch_set_ic_device(R.r.get(), gT.p, gp.p, gY.p);
      if (rc >= 0) rc = ct_batch_integrate(R.r.get(), b->t_stop);
      if (rc <= CT_ERR_INVALID) { R.rc = rc; R.err = g_err; return; }
      R.st.resize(m);
      if (cudaMemcpyAsync(R.st.data(), R.r->status.p, m * sizeof(int), cudaMemcpyDeviceToHost, R.r->stream) != cudaSuccess ||
          cudaStreamSynchronize(R.r->stream) != cudaSuccess) { R.rc = CT_ERR_CUDA; R.err = "status copy"; }
    };
    if (nconc > 1) {
      std::vector<std::thread> th;
      for (int c = 0; c < nconc; ++c) th.emplace_back(one, c);
      for (auto& t : th) t.join();
    } else one(0);
    for (int c = 0; c < nconc; ++c) if (runs[c].rc < 0) { for (auto& R : runs) if (R.s) cudaStreamDestroy(R.s); return fail(runs[c].rc, runs[c].err); }
    // every reactor takes the first rung that got through: masks on the host, merges on the device
    std::vector<int> taken(m, 0), still;
    DevBuf<int> ok;
    CT_CUDA(ok.alloc(m));
    for (int c = 0; c < nconc; ++c) {
      ct_batch* r = runs[c].r.get();
      std::vector<int> mask(m, -1);
      for (int jx = 0; jx < m; ++jx) if (!taken[jx] && runs[c].st[jx] >= 0) { mask[jx] = 0; taken[jx] = 1; }
      CT_CUDA(cudaMemcpyAsync(ok.p, mask.data(), m * sizeof(int), cudaMemcpyHostToDevice, b->stream));
      k_scatter_rows<double><<<blocks((int64_t)m * N), thr, 0, b->stream>>>(r->state.p, idx.p, ok.p, m, N, b->state.p);
      k_scatter_rows<double><<<blocks(m), thr, 0, b->stream>>>(r->tau.p, idx.p, ok.p, m, 1, b->tau.p);
      k_scatter_rows<double><<<blocks(m), thr, 0, b->stream>>>(r->tcur.p, idx.p, ok.p, m, 1, b->tcur.p);
      k_scatter_rows<long long><<<blocks((int64_t)m * 8), thr, 0, b->stream>>>(r->stats.p, idx.p, ok.p, m, 8, b->stats.p);
      k_scatter_rows<double><<<blocks((int64_t)m * 6), thr, 0, b->stream>>>(r->stats_ex.p, idx.p, ok.p, m, 6, b->stats_ex.p);
      k_scatter_rows<int><<<blocks(m), thr, 0, b->stream>>>(r->status.p, idx.p, ok.p, m, 1, b->status.p);
      CT_CUDA(cudaGetLastError());
      CT_CUDA(cudaStreamSynchronize(b->stream));  // `mask` goes out of scope
      b->launches += r->launches; b->aux_launches += r->aux_launches + 6;
    }
    for (auto& R : runs) if (R.s) cudaStreamDestroy(R.s);
    for (int jx = 0; jx < m; ++jx) if (!taken[jx]) still.push_back(bad[jx]);
    bad.swap(still);
    rung += nconc;
  }
  return 0;
}

}  // namespace

extern "C" {

// Integrate to the stop time, then the recovery ladder over the reactors that ended with a negative CVODE status (what
// IStrategy::Integrate only sketches, interface.cpp:462-473).  Fresh batches only.
int ct_batch_integrate_with_retry(ct_batch* b, int64_t* n_failed_first, int64_t* n_failed_final) {
  if (!b) return fail(CT_ERR_INVALID, "null batch");
  if (!b->fresh) return fail(CT_ILL_INPUT, "IntegrateWithRetry needs a fresh batch (set the initial conditions again)");
  int rc = ct_batch_integrate(b, b->t_stop);
  if (rc <= CT_ERR_INVALID) return rc;
  std::vector<int> st(b->n);
  CT_CUDA(cudaMemcpy(st.data(), b->status.p, b->n * sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<int> bad;
  for (int64_t i = 0; i < b->n; ++i) if (st[i] < 0) bad.push_back((int)i);
  if (n_failed_first) *n_failed_first = (int64_t)bad.size();
  if (int e = retry_ladder(b, bad)) return e;
  if (n_failed_final) *n_failed_final = (int64_t)bad.size();
  int out = 0;
  if (int e = summarize_status(b, &out)) return e;
  return out;
}

int ct_batch_integrate_trajectory(ct_batch* b, int n_out, double* traj) {
  if (int r = precheck_integrate(b)) return r;
  if (n_out < 1 || !traj) return fail(CT_ERR_INVALID, "bad trajectory arguments");
  if (!b->fresh) return fail(CT_ILL_INPUT, "trajectory integration needs a fresh batch (set the initial conditions again)");
  if (b->n == 0) return CT_SUCCESS;
  CT_CUDA(b->traj.alloc((size_t)b->n * n_out * b->N));
  BatchArgs a = make_args(b);
  a.n_out = n_out; a.traj = b->traj.p;
  CT_CUDA(cudaMemsetAsync(b->status.p, 0, b->n * sizeof(int), b->stream));
  if (int r = launch_integrate(b, a)) return r;
  b->fresh = false;
  CT_CUDA(cudaMemcpy(traj, b->traj.p, (size_t)b->n * n_out * b->N * sizeof(double), cudaMemcpyDeviceToHost));
  int out = 0;
  if (int r = summarize_status(b, &out)) return r;
  return out;
}

int ct_batch_get_state(ct_batch* b, double* out) {
  if (!b || !out) return fail(CT_ERR_INVALID, "null argument");
  if (!b->have_ic) return fail(CT_ILL_INPUT, "initial conditions not set");
  CT_CUDA(cudaMemcpyAsync(out, b->state.p, (size_t)b->n * b->N * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  CT_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}
int ct_batch_get_state_device(ct_batch* b, const double** dptr) {
  if (!b || !dptr) return fail(CT_ERR_INVALID, "null argument");
  *dptr = b->state.p;
  return 0;
}
#define CT_GETTER(name, buf, type, mult)                                                              \
  int name(ct_batch* b, type* out) {                                                                 \
    if (!b || !out) return fail(CT_ERR_INVALID, "null argument");                                     \
    if (!b->buf.p || b->launches == 0) return fail(CT_ILL_INPUT, "nothing integrated yet");           \
    CT_CUDA(cudaMemcpyAsync(out, b->buf.p, (size_t)b->n * (mult) * sizeof(type), cudaMemcpyDeviceToHost, b->stream)); \
    CT_CUDA(cudaStreamSynchronize(b->stream));                                                        \
    return 0;                                                                                         \
  }
CT_GETTER(ct_batch_get_time, tcur, double, 1)
CT_GETTER(ct_batch_get_ignition, tau, double, 1)
CT_GETTER(ct_batch_get_status, status, int32_t, 1)
int ct_batch_get_stats(ct_batch* b, int64_t* out) {
  if (!b || !out) return fail(CT_ERR_INVALID, "null argument");
  if (!b->stats.p || b->launches == 0) return fail(CT_ILL_INPUT, "nothing integrated yet");
  CT_CUDA(cudaMemcpyAsync(out, b->stats.p, (size_t)b->n * 8 * sizeof(long long), cudaMemcpyDeviceToHost, b->stream));
  CT_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}
int ct_batch_get_stats_ex(ct_batch* b, double* out) {
  if (!b || !out) return fail(CT_ERR_INVALID, "null argument");
  if (!b->stats_ex.p || b->launches == 0) return fail(CT_ILL_INPUT, "nothing integrated yet");
  CT_CUDA(cudaMemcpyAsync(out, b->stats_ex.p, (size_t)b->n * 6 * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  CT_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}
double ct_batch_last_kernel_ms(const ct_batch* b) { return b ? b->last_ms : 0.0; }
int64_t ct_batch_launch_count(const ct_batch* b) { return b ? b->launches + b->aux_launches : 0; }
int ct_batch_warps_per_sm(const ct_batch* b) { return b ? b->wpb : 0; }

// ---------------------------------------------------------------- stand-alone kernels
static int standalone_cfg(ct_batch* b, int* wpb, int* grid, size_t* smem, bool need_scratch) {
  if (int r = configure_launch(b)) return r;
  {
    DeviceInfo di0;
    if (int r = device_info(di0)) return r;
    const size_t slab_c = (size_t)make_layout(b->N, b->mech->desc.R).doubles * sizeof(double);
    *wpb = (int)std::min<size_t>(8, ((size_t)di0.max_smem_optin - CT_HDR_BYTES - b->mech->desc.blob_bytes) / slab_c);
  }
  const WarpLayout L = make_layout(b->N, b->mech->desc.R);
  *smem = CT_HDR_BYTES + b->mech->desc.blob_bytes + (size_t)*wpb * L.doubles * sizeof(double);
  DeviceInfo di;
  if (int r = device_info(di)) return r;
  *grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)di.sms * 1, (b->n + *wpb - 1) / *wpb));
  if (need_scratch) CT_CUDA(b->scratch.alloc(std::max<size_t>(b->scratch.n, (size_t)*grid * *wpb * warp_scratch_doubles(b->N, b->mech->desc.R))));
  return 0;
}

int ct_rhs_eval(ct_batch* b, const double* states, const double* times, double* ydot, double* wdot, double* qf, double* qr) {
  if (!b || !states || !ydot) return fail(CT_ERR_INVALID, "null argument");
  if (!b->have_ic) return fail(CT_ILL_INPUT, "initial conditions not set");
  if ((qf == nullptr) != (qr == nullptr)) return fail(CT_ERR_INVALID, "qf and qr must be given together");
  const int R = b->mech->desc.R;
  int wpb, grid; size_t smem;
  if (int r = standalone_cfg(b, &wpb, &grid, &smem, false)) return r;
  const unsigned char* blob; const int* perm;
  if (int r = b->mech->on_device(&blob, &perm)) return r;
  const size_t nN = (size_t)b->n * b->N;
  CT_CUDA(b->tmp_in.alloc(nN)); CT_CUDA(b->tmp_out.alloc(nN)); CT_CUDA(b->tmp_w.alloc((size_t)b->n * b->K));
  CT_CUDA(cudaMemcpyAsync(b->tmp_in.p, states, nN * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  if (times) { CT_CUDA(b->tmp_t.alloc(b->n)); CT_CUDA(cudaMemcpyAsync(b->tmp_t.p, times, b->n * sizeof(double), cudaMemcpyHostToDevice, b->stream)); }
  if (qf) { CT_CUDA(b->tmp_qf.alloc((size_t)b->n * R)); CT_CUDA(b->tmp_qr.alloc((size_t)b->n * R)); }
  BatchArgs a = make_args(b);
  a.pool_n = 0;
  if (int r = set_smem(k_rhs, smem)) return r;
  k_rhs<<<grid, wpb * 32, smem, b->stream>>>(b->mech->desc, blob, a, b->tmp_in.p, times ? b->tmp_t.p : nullptr, b->tmp_out.p,
                                             b->tmp_w.p, qf ? b->tmp_qf.p : nullptr, qf ? b->tmp_qr.p : nullptr, perm);
  CT_CUDA(cudaGetLastError());
  b->launches += 1;
  CT_CUDA(cudaMemcpyAsync(ydot, b->tmp_out.p, nN * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  if (wdot) CT_CUDA(cudaMemcpyAsync(wdot, b->tmp_w.p, (size_t)b->n * b->K * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  if (qf) {
    CT_CUDA(cudaMemcpyAsync(qf, b->tmp_qf.p, (size_t)b->n * R * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    CT_CUDA(cudaMemcpyAsync(qr, b->tmp_qr.p, (size_t)b->n * R * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  }
  CT_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}

int ct_rhs_time(ct_batch* b, const double* states, int reps, double* ms) {
  if (!b || !states || !ms || reps < 1) return fail(CT_ERR_INVALID, "bad argument");
  if (!b->have_ic) return fail(CT_ILL_INPUT, "initial conditions not set");
  int wpb, grid; size_t smem;
  if (int r = standalone_cfg(b, &wpb, &grid, &smem, false)) return r;
  const unsigned char* blob; const int* perm;
  if (int r = b->mech->on_device(&blob, &perm)) return r;
  const size_t nN = (size_t)b->n * b->N;
  CT_CUDA(b->tmp_in.alloc(nN)); CT_CUDA(b->tmp_out.alloc(nN)); CT_CUDA(b->tmp_w.alloc((size_t)b->n * b->K));
  CT_CUDA(cudaMemcpy(b->tmp_in.p, states, nN * sizeof(double), cudaMemcpyHostToDevice));
  if (!b->ev0) { CT_CUDA(cudaEventCreate(&b->ev0)); CT_CUDA(cudaEventCreate(&b->ev1)); }
  BatchArgs a = make_args(b);
  a.pool_n = 0;
  if (int r = set_smem(k_rhs, smem)) return r;
  for (int w = 0; w < 3; ++w) k_rhs<<<grid, wpb * 32, smem, b->stream>>>(b->mech->desc, blob, a, b->tmp_in.p, nullptr, b->tmp_out.p, b->tmp_w.p, nullptr, nullptr, perm);
  CT_CUDA(cudaEventRecord(b->ev0, b->stream));
  for (int w = 0; w < reps; ++w) k_rhs<<<grid, wpb * 32, smem, b->stream>>>(b->mech->desc, blob, a, b->tmp_in.p, nullptr, b->tmp_out.p, b->tmp_w.p, nullptr, nullptr, perm);
  CT_CUDA(cudaEventRecord(b->ev1, b->stream));
  CT_CUDA(cudaStreamSynchronize(b->stream));
  CT_CUDA(cudaGetLastError());
  float t = 0.f;
  CT_CUDA(cudaEventElapsedTime(&t, b->ev0, b->ev1));
  *ms = t / reps;
  b->launches += reps + 3;
  return 0;
}

// occupancy scan of K1 for the roofline report: the right-hand side in the fused kernel's slab layout (no Newton matrix in
// the per-warp slab), `warps_per_block` warps per CTA and `blocks_per_sm` CTAs per SM under the matching register cap
int ct_rhs_time_cfg(ct_batch* b, const double* states, int reps, int warps_per_block, int blocks_per_sm, double* ms) {
  if (!b || !states || !ms || reps < 1 || warps_per_block < 1 || warps_per_block > 32 || blocks_per_sm < 1) return fail(CT_ERR_INVALID, "bad argument");
  if (!b->have_ic) return fail(CT_ILL_INPUT, "initial conditions not set");
  DeviceInfo di;
  if (int r = device_info(di)) return r;
  const unsigned char* blob; const int* perm;
  if (int r = b->mech->on_device(&blob, &perm)) return r;
  const WarpLayout L = make_layout(b->N, b->mech->desc.R, false);
  const size_t smem = CT_HDR_BYTES + b->mech->desc.blob_bytes + (size_t)warps_per_block * L.doubles * sizeof(double);
  if (smem > (size_t)di.max_smem_optin) return fail(CT_ERR_UNSUPPORTED, "too many warps per block for one SM's shared memory");
  const size_t nN = (size_t)b->n * b->N;
  CT_CUDA(b->tmp_in.alloc(nN)); CT_CUDA(b->tmp_out.alloc(nN)); CT_CUDA(b->tmp_w.alloc((size_t)b->n * b->K));
  CT_CUDA(cudaMemcpy(b->tmp_in.p, states, nN * sizeof(double), cudaMemcpyHostToDevice));
  if (!b->ev0) { CT_CUDA(cudaEventCreate(&b->ev0)); CT_CUDA(cudaEventCreate(&b->ev1)); }
  BatchArgs a = make_args(b);
  a.pool_n = 1;  // slab without the matrix
  const int threads = warps_per_block * 32, grid = di.sms * blocks_per_sm;
  const bool st30 = b->use_static_shape && b->mech->shape == 1;
#define CT_RHS_ARGS b->mech->desc, blob, a, b->tmp_in.p, nullptr, b->tmp_out.p, b->tmp_w.p, nullptr, nullptr, perm
  auto launch = [&]() {
    if (st30) {
      if (threads <= 384) k_rhs_t<384, ShapeGri30><<<grid, threads, smem, b->stream>>>(CT_RHS_ARGS);
      else if (threads <= 512) k_rhs_t<512, ShapeGri30><<<grid, threads, smem, b->stream>>>(CT_RHS_ARGS);
      else k_rhs_t<768, ShapeGri30><<<grid, threads, smem, b->stream>>>(CT_RHS_ARGS);
    } else {
      if (threads <= 384) k_rhs_t<384, ShapeRt><<<grid, threads, smem, b->stream>>>(CT_RHS_ARGS);
      else if (threads <= 512) k_rhs_t<512, ShapeRt><<<grid, threads, smem, b->stream>>>(CT_RHS_ARGS);
      else k_rhs_t<768, ShapeRt><<<grid, threads, smem, b->stream>>>(CT_RHS_ARGS);
    }
  };
#undef CT_RHS_ARGS
  if (threads > 768) return fail(CT_ERR_INVALID, "at most 24 warps per block");
  if (int r = set_smem(k_rhs_t<384, ShapeGri30>, smem)) return r;
  if (int r = set_smem(k_rhs_t<512, ShapeGri30>, smem)) return r;
  if (int r = set_smem(k_rhs_t<768, ShapeGri30>, smem)) return r;
  if (int r = set_smem(k_rhs_t<384, ShapeRt>, smem)) return r;
  if (int r = set_smem(k_rhs_t<512, ShapeRt>, smem)) return r;
  if (int r = set_smem(k_rhs_t<768, ShapeRt>, smem)) return r;
  for (int w = 0; w < 3; ++w) launch();
  CT_CUDA(cudaGetLastError());
  CT_CUDA(cudaEventRecord(b->ev0, b->stream));
  for (int w = 0; w < reps; ++w) launch();
  CT_CUDA(cudaEventRecord(b->ev1, b->stream));
  CT_CUDA(cudaStreamSynchronize(b->stream));
  CT_CUDA(cudaGetLastError());
  float t = 0.f;
  CT_CUDA(cudaEventElapsedTime(&t, b->ev0, b->ev1));
  *ms = t / reps;
  b->launches += reps + 3;
  return 0;
}

static int jac_launch(ct_batch* b, int mode, const double* d_states, const double* d_times, double* d_J, int reps, double* ms) {
  int wpb, grid; size_t smem;
  if (int r = standalone_cfg(b, &wpb, &grid, &smem, true)) return r;
  const unsigned char* blob;
  if (int r = b->mech->on_device(&blob, nullptr)) return r;
  if (!b->ev0) { CT_CUDA(cudaEventCreate(&b->ev0)); CT_CUDA(cudaEventCreate(&b->ev1)); }
  BatchArgs a = make_args(b);
  a.pool_n = 0;
  a.jac_mode = mode;
  if (int r = set_smem(k_jacobian, smem)) return r;
  if (ms) for (int w = 0; w < 2; ++w) k_jacobian<<<grid, wpb * 32, smem, b->stream>>>(b->mech->desc, blob, a, d_states, d_times, 1e-8, d_J);
  CT_CUDA(cudaEventRecord(b->ev0, b->stream));
  for (int w = 0; w < reps; ++w) k_jacobian<<<grid, wpb * 32, smem, b->stream>>>(b->mech->desc, blob, a, d_states, d_times, 1e-8, d_J);
  CT_CUDA(cudaEventRecord(b->ev1, b->stream));
  CT_CUDA(cudaStreamSynchronize(b->stream));
  CT_CUDA(cudaGetLastError());
  float t = 0.f;
  CT_CUDA(cudaEventElapsedTime(&t, b->ev0, b->ev1));
  if (ms) *ms = t / reps;
  b->launches += reps;
  return 0;
}

int ct_jacobian_eval(ct_batch* b, int mode, const double* states, const double* times, double* J) {
  if (!b || !states || !J) return fail(CT_ERR_INVALID, "null argument");
  if (!b->have_ic) return fail(CT_ILL_INPUT, "initial conditions not set");
  const size_t nN = (size_t)b->n * b->N;
  CT_CUDA(b->tmp_in.alloc(nN)); CT_CUDA(b->tmp_out.alloc(nN * b->N));
  CT_CUDA(cudaMemcpy(b->tmp_in.p, states, nN * sizeof(double), cudaMemcpyHostToDevice));
  if (times) { CT_CUDA(b->tmp_t.alloc(b->n)); CT_CUDA(cudaMemcpy(b->tmp_t.p, times, b->n * sizeof(double), cudaMemcpyHostToDevice)); }
  if (int r = jac_launch(b, mode, b->tmp_in.p, times ? b->tmp_t.p : nullptr, b->tmp_out.p, 1, nullptr)) return r;
  CT_CUDA(cudaMemcpy(J, b->tmp_out.p, nN * b->N * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}
int ct_jacobian_time(ct_batch* b, int mode, const double* states, int reps, double* ms) {
  if (!b || !states || !ms || reps < 1) return fail(CT_ERR_INVALID, "bad argument");
  if (!b->have_ic) return fail(CT_ILL_INPUT, "initial conditions not set");
  const size_t nN = (size_t)b->n * b->N;
  CT_CUDA(b->tmp_in.alloc(nN)); CT_CUDA(b->tmp_out.alloc(nN * b->N));
  CT_CUDA(cudaMemcpy(b->tmp_in.p, states, nN * sizeof(double), cudaMemcpyHostToDevice));
  return jac_launch(b, mode, b->tmp_in.p, nullptr, b->tmp_out.p, reps, ms);
}

static int lu_run(int N, int64_t n, const double* dA, const double* db, double* dx, int* dinfo, int reps, int solves, double* ms, int linsol) {
  DeviceInfo di;
  if (int r = device_info(di)) return r;
  const WarpLayout L = make_layout(N, 4);
  const size_t slab = (size_t)L.doubles * sizeof(double);
  int wpb = (int)std::min<size_t>(8, (size_t)di.max_smem_optin / slab);
  if (wpb < 1) return fail(CT_ERR_UNSUPPORTED, "matrix too large for shared memory");
  const size_t smem = wpb * slab;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(di.sms, (n + wpb - 1) / wpb));
  if (int r = set_smem(k_lu_solve, smem)) return r;
  cudaEvent_t e0, e1;
  CT_CUDA(cudaEventCreate(&e0)); CT_CUDA(cudaEventCreate(&e1));
  if (ms) for (int w = 0; w < 2; ++w) k_lu_solve<<<grid, wpb * 32, smem>>>(N, n, dA, db, dx, dinfo, solves, linsol);
  CT_CUDA(cudaEventRecord(e0));
  for (int w = 0; w < reps; ++w) k_lu_solve<<<grid, wpb * 32, smem>>>(N, n, dA, db, dx, dinfo, solves, linsol);
  CT_CUDA(cudaEventRecord(e1));
  CT_CUDA(cudaDeviceSynchronize());
  CT_CUDA(cudaGetLastError());
  float t = 0.f;
  CT_CUDA(cudaEventElapsedTime(&t, e0, e1));
  if (ms) *ms = t / reps;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}

int ct_lu_solve(int N, int64_t n, const double* A, const double* b, double* x, int32_t* info, int linsol) {
  if (int r = require_device()) return r;
  if (N < 1 || N > CT_NP || n < 0 || !A || !b || !x) return fail(CT_ERR_INVALID, "bad argument");
  DevBuf<double> dA, db, dx; DevBuf<int> dinfo;
  CT_CUDA(dA.alloc((size_t)n * N * N)); CT_CUDA(db.alloc((size_t)n * N)); CT_CUDA(dx.alloc((size_t)n * N)); CT_CUDA(dinfo.alloc(n));
  CT_CUDA(cudaMemcpy(dA.p, A, (size_t)n * N * N * sizeof(double), cudaMemcpyHostToDevice));
  CT_CUDA(cudaMemcpy(db.p, b, (size_t)n * N * sizeof(double), cudaMemcpyHostToDevice));
  CT_CUDA(cudaMemset(dx.p, 0, (size_t)n * N * sizeof(double)));
  if (n > 0) if (int r = lu_run(N, n, dA.p, db.p, dx.p, dinfo.p, 1, 1, nullptr, linsol)) return r;
  CT_CUDA(cudaMemcpy(x, dx.p, (size_t)n * N * sizeof(double), cudaMemcpyDeviceToHost));
  if (info) CT_CUDA(cudaMemcpy(info, dinfo.p, n * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}
int ct_lu_time(int N, int64_t n, int reps, int solves, int linsol, double* ms) {
  if (int r = require_device()) return r;
  if (N < 1 || N > CT_NP || n < 1 || reps < 1 || !ms) return fail(CT_ERR_INVALID, "bad argument");
  std::vector<double> A((size_t)n * N * N), b((size_t)n * N, 1.0);
  unsigned long long s = 88172645463325252ULL;
  for (auto& v : A) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (double)(s >> 11) / 9007199254740992.0 - 0.5; }
  for (int64_t i = 0; i < n; ++i) for (int d = 0; d < N; ++d) A[(size_t)i * N * N + d * N + d] += 4.0;
  DevBuf<double> dA, db, dx; DevBuf<int> dinfo;
  CT_CUDA(dA.alloc(A.size())); CT_CUDA(db.alloc(b.size())); CT_CUDA(dx.alloc(b.size())); CT_CUDA(dinfo.alloc(n));
  CT_CUDA(cudaMemcpy(dA.p, A.data(), A.size() * sizeof(double), cudaMemcpyHostToDevice));
  CT_CUDA(cudaMemcpy(db.p, b.data(), b.size() * sizeof(double), cudaMemcpyHostToDevice));
  return lu_run(N, n, dA.p, db.p, dx.p, dinfo.p, reps, solves, ms, linsol);
}

int ct_measure_fp64_peak(double* tflops) {
  if (int r = require_device()) return r;
  if (!tflops) return fail(CT_ERR_INVALID, "null argument");
  DeviceInfo di;
  if (int r = device_info(di)) return r;
  const int blocks = di.sms * 8, threads = 256, iters = 8192;
  DevBuf<double> out;
  CT_CUDA(out.alloc((size_t)blocks * threads));
  cudaEvent_t e0, e1;
  CT_CUDA(cudaEventCreate(&e0)); CT_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CT_CUDA(cudaEventRecord(e0));
    k_dfma_peak<<<blocks, threads>>>(out.p, iters);
    CT_CUDA(cudaEventRecord(e1));
    CT_CUDA(cudaDeviceSynchronize());
    float t = 0.f;
    CT_CUDA(cudaEventElapsedTime(&t, e0, e1));
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    best = std::max(best, fl / (t * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops = best;
  return 0;
}

int ct_roberts_selftest(int n_out, const double* t_out, double rtol, const double* atol3, double* y_out, int64_t* stats8) {
  if (int r = require_device()) return r;
  if (n_out < 1 || !t_out || !atol3 || !y_out) return fail(CT_ERR_INVALID, "bad argument");
  ct_batch b;
  b.mech = nullptr; b.roberts = true; b.rt = 0; b.N = 3; b.K = 0; b.n = 1;
  b.rtol = rtol; b.atol = 0.0; b.mxstep = 500; b.jac_mode = CT_JAC_ANALYTIC;
  b.t_stop = 1e300;
  CT_CUDA(b.atol_vec.alloc(3));
  CT_CUDA(cudaMemcpy(b.atol_vec.p, atol3, 3 * sizeof(double), cudaMemcpyHostToDevice));
  CT_CUDA(b.state.alloc(3));
  const double y0[3] = {1.0, 0.0, 0.0};
  CT_CUDA(cudaMemcpy(b.state.p, y0, sizeof y0, cudaMemcpyHostToDevice));
  b.have_ic = true;
  if (int r = ensure_runtime_buffers(&b)) return r;
  BatchArgs a = make_args(&b);
  DevBuf<double> d_tout;
  CT_CUDA(d_tout.alloc(n_out));
  CT_CUDA(cudaMemcpy(d_tout.p, t_out, n_out * sizeof(double), cudaMemcpyHostToDevice));
  CT_CUDA(b.traj.alloc((size_t)n_out * 3));
  CT_CUDA(cudaMemset(b.status.p, 0, sizeof(int)));
  a.n_out = n_out; a.traj = b.traj.p; a.out_times = d_tout.p; a.T0 = nullptr; a.p0 = nullptr; a.rho0 = nullptr;
  if (int r = launch_integrate(&b, a)) return r;
  int st = 0;
  CT_CUDA(cudaMemcpy(&st, b.status.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (st < 0) return fail(st, "device stepper failed on the Roberts problem");
  CT_CUDA(cudaMemcpy(y_out, b.traj.p, (size_t)n_out * 3 * sizeof(double), cudaMemcpyDeviceToHost));
  if (stats8) CT_CUDA(cudaMemcpy(stats8, b.stats.p, 8 * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"

#include "ensemble.inl"
