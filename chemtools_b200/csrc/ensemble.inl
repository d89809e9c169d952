// Multi-GPU ensemble driver (included at the end of capi.cu).
//
// The reference integrates one reactor at a time on one host thread (Base::Solve, src/applications/reactors/base.cpp:
// 123-170) and has no multi-device path; SURVEY.md section 8(e) defines this one: reactors are independent closed systems,
// so the index range is partitioned over the GPUs of one box, one host thread + one CUDA stream per GPU, no collective on
// the solve path, and every GPU copies its results with plain cudaMemcpyAsync D2H into disjoint slices of the caller's
// (ideally pinned, ct_host_alloc) host arrays, which the results writer consumes (src/results/hdf5_writer.cpp:101-110).
// Each shard is processed in chunks of `chunk` reactors so that the reactor-indexed scratch of the time-sliced work ring
// stays below its 4 GiB bound and device memory use does not depend on the ensemble size.
//
// A shard is nothing but a sequence of ordinary batches (ct_batch_create / set_ic / integrate, results read from the batch's
// device arrays) followed by the library's own recovery ladder (retry_ladder in capi.cu), which is why a sharded run is
// bit-identical to a single-GPU run of the same reactors.

struct ct_ensemble {
  ct_mech* mech = nullptr;
  int rt = 0, N = 0, K = 0;
  int64_t n = 0, chunk = 65536;
  std::vector<int> devices;
  int partition = CT_PARTITION_BLOCK;
  double rtol = 1e-8, atol = 1e-14, t_stop = 1e-3;
  int64_t mxstep = 500, first_pass = 10000;
  int retry = 1, hot_first = 1, streams = 2;
  ct_batch_configure_fn configure = nullptr;
  void* configure_user = nullptr;
  int nt = 0;
  const double *tab_t = nullptr, *tab_T = nullptr, *tab_p = nullptr;  // borrowed until integrate returns
  struct Shard {
    int64_t count = 0, launches = 0, failed_first = 0, failed_final = 0, chunks = 0;
    double kernel_ms = 0.0, wall_s = 0.0;
    int rc = 0, worst = 0;
    bool all_tstop = true;
    std::string err;
    std::vector<int64_t> bad;  // global indices of the reactors that ended with a negative status
    std::mutex mu;
    Shard() = default;
    Shard(const Shard& o) : count(o.count), launches(o.launches), failed_first(o.failed_first), failed_final(o.failed_final), chunks(o.chunks),
                            kernel_ms(o.kernel_ms), wall_s(o.wall_s), rc(o.rc), worst(o.worst), all_tstop(o.all_tstop), err(o.err), bad(o.bad) {}
    Shard& operator=(const Shard& o) {
      count = o.count; launches = o.launches; failed_first = o.failed_first; failed_final = o.failed_final; chunks = o.chunks;
      kernel_ms = o.kernel_ms; wall_s = o.wall_s; rc = o.rc; worst = o.worst; all_tstop = o.all_tstop; err = o.err; bad = o.bad;
      return *this;
    }
  };
  std::vector<Shard> shards;
  int64_t shard_count(int g) const {
    const int64_t G = (int64_t)devices.size();
    if (partition == CT_PARTITION_CYCLIC) return (n - g + G - 1) / G;
    return (n * (g + 1) + G - 1) / G - (n * g + G - 1) / G;
  }
  // step budget of the chunk launches: a launch lasts as long as its slowest reactor (one warp, 40-60 us per step), so the
  // few reactors that need more than `first_pass` steps are cut off there and integrated again, with the full budget, in
  // one batch per shard after its last chunk (one long tail per shard instead of one per chunk)
  int64_t first_pass_budget() const { return (retry && first_pass > 0) ? std::min(first_pass, mxstep) : mxstep; }
  // global index of the j-th reactor of shard g
  int64_t global_index(int g, int64_t j) const {
    const int64_t G = (int64_t)devices.size();
    if (partition == CT_PARTITION_CYCLIC) return g + j * G;
    return (n * g + G - 1) / G + j;
  }
};

namespace {

// T-p table or V-dV/dt table, whichever the reactor type takes
int batch_set_table(ct_batch* b, int nt, const double* t, const double* A, const double* B) {
  return b->rt == CT_TABLE_V ? ct_batch_set_v_table(b, nt, t, A, B) : ct_batch_set_tp_table(b, nt, t, A, B);
}

struct PinnedBuf {
  void* p = nullptr;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  cudaError_t alloc(size_t bytes) { return cudaHostAlloc(&p, std::max<size_t>(bytes, 8), cudaHostAllocPortable); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

int configure_batch(ct_ensemble* e, ct_batch* b, cudaStream_t stream, bool first_pass) {
  b->stream = stream;
  if (int r = ct_batch_set_tolerances(b, e->rtol, e->atol)) return r;
  if (int r = ct_batch_set_stop_time(b, e->t_stop)) return r;
  if (int r = ct_batch_set_max_steps(b, first_pass ? e->first_pass_budget() : e->mxstep)) return r;
  if (e->configure) {
    const int r = e->configure(b, e->configure_user);
    if (r < 0) return r <= CT_ERR_INVALID ? r : fail(CT_ERR_INVALID, "configure callback failed");
  }
  return 0;
}

// one worker = one host thread, one stream on the shard's device, every `stride`-th chunk of the shard as a sequence of
// ordinary batches.  Two workers per device keep two launches in flight: the CTAs of a launch leave as soon as the work
// ring is empty, so the next chunk's CTAs take over the SMs while the last (slowest) reactors of this one finish.
int run_worker(ct_ensemble* e, int g, int w, int stride, const double* T, const double* p, const double* Y, double* state, double* tau,
               int32_t* status, int64_t* stats) {
  ct_ensemble::Shard& S = e->shards[g];
  CT_CUDA(cudaSetDevice(e->devices[g]));
  cudaStream_t stream = nullptr;
  CT_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{stream};
  const int K = e->K, N = e->N, nt = e->nt;
  const bool cyclic = e->partition == CT_PARTITION_CYCLIC && e->devices.size() > 1;
  const int64_t cnt = S.count, chunk = std::max<int64_t>(1, std::min<int64_t>(e->chunk, cnt));
  // cyclic partition: strided gather / scatter through pinned staging buffers; block partition: the caller's arrays
  // are read and written in place (disjoint contiguous slices)
  PinnedBuf sT, sp, sY, sState, sTau, sStatus, sStats, sTabT, sTabp;
  if (cyclic) {
    CT_CUDA(sT.alloc(chunk * 8)); CT_CUDA(sp.alloc(chunk * 8)); CT_CUDA(sY.alloc((size_t)chunk * K * 8));
    if (state) CT_CUDA(sState.alloc((size_t)chunk * N * 8));
    if (tau) CT_CUDA(sTau.alloc(chunk * 8));
    if (status) CT_CUDA(sStatus.alloc(chunk * 4));
    if (stats) CT_CUDA(sStats.alloc((size_t)chunk * 64));
    if (nt) { CT_CUDA(sTabT.alloc((size_t)chunk * nt * 8)); CT_CUDA(sTabp.alloc((size_t)chunk * nt * 8)); }
  }
  std::unique_ptr<ct_batch> batch;
  std::vector<int32_t> order, st_host;
  std::vector<int64_t> bad;
  int64_t launches = 0, chunks = 0;
  double kernel_ms = 0.0;
  int worst = 0;
  bool all_tstop = true;
  for (int64_t c0 = (int64_t)w * chunk; c0 < cnt; c0 += (int64_t)stride * chunk) {
    const int64_t m = std::min<int64_t>(chunk, cnt - c0);
    if (!batch || batch->n != m) {
      int stc = 0;
      batch.reset(ct_batch_create(e->mech, e->rt, m, &stc));
      if (!batch) return stc;
      if (int r = configure_batch(e, batch.get(), stream, true)) return r;
    }
    const double *cT, *cp, *cY, *cTabT = nullptr, *cTabp = nullptr;
    const int64_t i0 = e->global_index(g, c0);
    if (cyclic) {
      double *dT = sT.as<double>(), *dp = sp.as<double>(), *dY = sY.as<double>();
      for (int64_t j = 0; j < m; ++j) {
        const int64_t i = e->global_index(g, c0 + j);
        dT[j] = T[i]; dp[j] = p[i];
        std::memcpy(dY + j * K, Y + i * K, (size_t)K * 8);
        if (nt) {
          std::memcpy(sTabT.as<double>() + j * nt, e->tab_T + i * nt, (size_t)nt * 8);
          std::memcpy(sTabp.as<double>() + j * nt, e->tab_p + i * nt, (size_t)nt * 8);
        }
      }
      cT = dT; cp = dp; cY = dY; cTabT = sTabT.as<double>(); cTabp = sTabp.as<double>();
    } else {
      cT = T + i0; cp = p + i0; cY = Y + i0 * K;
      if (nt) { cTabT = e->tab_T + i0 * nt; cTabp = e->tab_p + i0 * nt; }
    }
    const int64_t launches_before = batch->launches + batch->aux_launches;
    if (int r = ct_batch_set_ic(batch.get(), cT, cp, cY)) return r;
    if (nt) { if (int r = batch_set_table(batch.get(), nt, e->tab_t, cTabT, cTabp)) return r; }
    if (e->hot_first && m > 1) {
      // expected-cost order: hottest first, so that the longest integrations start first (shortens the tail of the queue)
      order.resize(m);
      for (int64_t j = 0; j < m; ++j) order[j] = (int32_t)j;
      std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return cT[a] > cT[b]; });
      if (int r = ct_batch_set_order(batch.get(), order.data())) return r;
    }
    const int rc = ct_batch_integrate(batch.get(), e->t_stop);
    if (rc <= CT_ERR_INVALID) return rc;
    kernel_ms += batch->last_ms; chunks += 1;
    // results: plain cudaMemcpyAsync D2H on the worker's stream into the caller's slices (or the staging buffers)
    double* oState = state ? (cyclic ? sState.as<double>() : state + i0 * N) : nullptr;
    double* oTau = tau ? (cyclic ? sTau.as<double>() : tau + i0) : nullptr;
    int32_t* oStatus = status ? (cyclic ? sStatus.as<int32_t>() : status + i0) : nullptr;
    int64_t* oStats = stats ? (cyclic ? sStats.as<int64_t>() : stats + i0 * 8) : nullptr;
    if (oState) CT_CUDA(cudaMemcpyAsync(oState, batch->state.p, (size_t)m * N * 8, cudaMemcpyDeviceToHost, stream));
    if (oTau) CT_CUDA(cudaMemcpyAsync(oTau, batch->tau.p, (size_t)m * 8, cudaMemcpyDeviceToHost, stream));
    if (oStats) CT_CUDA(cudaMemcpyAsync(oStats, batch->stats.p, (size_t)m * 64, cudaMemcpyDeviceToHost, stream));
    st_host.resize(m);
    CT_CUDA(cudaMemcpyAsync(oStatus ? oStatus : st_host.data(), batch->status.p, (size_t)m * 4, cudaMemcpyDeviceToHost, stream));
    CT_CUDA(cudaStreamSynchronize(stream));
    const int32_t* sv = oStatus ? oStatus : st_host.data();
    for (int64_t j = 0; j < m; ++j) {
      if (sv[j] < worst) worst = sv[j];
      if (sv[j] != CT_TSTOP_RETURN) all_tstop = false;
      if (sv[j] < 0) bad.push_back(e->global_index(g, c0 + j));
    }
    if (cyclic) {
      for (int64_t j = 0; j < m; ++j) {
        const int64_t i = e->global_index(g, c0 + j);
        if (state) std::memcpy(state + i * N, oState + j * N, (size_t)N * 8);
        if (tau) tau[i] = oTau[j];
        if (status) status[i] = oStatus[j];
        if (stats) std::memcpy(stats + i * 8, oStats + j * 8, 64);
      }
    }
    launches += batch->launches + batch->aux_launches - launches_before;
  }
  std::lock_guard<std::mutex> lock(S.mu);
  S.launches += launches; S.chunks += chunks; S.kernel_ms += kernel_ms;
  S.worst = std::min(S.worst, worst); S.all_tstop = S.all_tstop && all_tstop;
  S.failed_first += (int64_t)bad.size();
  S.bad.insert(S.bad.end(), bad.begin(), bad.end());
  return 0;
}

// recovery pass of one shard, once, after all its chunks: the reactors that ended with a negative CVODE status form one
// small batch; first the caller's own options with the full step budget (if the chunk launches had a smaller one), then
// the library's recovery ladder (retry_ladder in capi.cu: Jacobian clip 1; LU instead of the inverse; difference-quotient
// Jacobian + LU); the successful ones overwrite their rows.  Same options, same ladder, same per-reactor arithmetic:
// identical to ct_batch_integrate_with_retry on a single batch with the full budget.
int run_ladder(ct_ensemble* e, int g, const double* T, const double* p, const double* Y, double* state, double* tau, int32_t* status,
               int64_t* stats) {
  ct_ensemble::Shard& S = e->shards[g];
  std::vector<int64_t> gidx = S.bad;
  std::sort(gidx.begin(), gidx.end());
  if (gidx.empty()) { S.failed_final = 0; return 0; }
  CT_CUDA(cudaSetDevice(e->devices[g]));
  cudaStream_t stream = nullptr;
  CT_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{stream};
  const int K = e->K, N = e->N, nt = e->nt;
  const int64_t m = (int64_t)gidx.size();
  int stc = 0;
  std::unique_ptr<ct_batch> r(ct_batch_create(e->mech, e->rt, m, &stc));
  if (!r) return stc;
  if (int rc = configure_batch(e, r.get(), stream, false)) return rc;
  std::vector<double> hT(m), hp(m), hY((size_t)m * K), hTabT, hTabp;
  if (nt) { hTabT.resize((size_t)m * nt); hTabp.resize((size_t)m * nt); }
  for (int64_t j = 0; j < m; ++j) {
    const int64_t i = gidx[j];
    hT[j] = T[i]; hp[j] = p[i];
    std::memcpy(hY.data() + j * K, Y + i * K, (size_t)K * 8);
    if (nt) {
      std::memcpy(hTabT.data() + j * nt, e->tab_T + i * nt, (size_t)nt * 8);
      std::memcpy(hTabp.data() + j * nt, e->tab_p + i * nt, (size_t)nt * 8);
    }
  }
  if (int rc = ct_batch_set_ic(r.get(), hT.data(), hp.data(), hY.data())) return rc;
  if (nt) { if (int rc = batch_set_table(r.get(), nt, e->tab_t, hTabT.data(), hTabp.data())) return rc; }
  std::vector<int> bad;
  std::vector<int32_t> oStatus(m);
  if (e->first_pass_budget() < e->mxstep) {
    const int rc = ct_batch_integrate(r.get(), e->t_stop);
    if (rc <= CT_ERR_INVALID) return rc;
    CT_CUDA(cudaMemcpyAsync(oStatus.data(), r->status.p, (size_t)m * 4, cudaMemcpyDeviceToHost, stream));
    CT_CUDA(cudaStreamSynchronize(stream));
    for (int64_t j = 0; j < m; ++j) if (oStatus[j] < 0) bad.push_back((int)j);
  } else {  // the chunk launches already had the full budget: straight to the ladder
    if (int rc = precheck_integrate(r.get())) return rc;
    CT_CUDA(cudaMemsetAsync(r->status.p, 0xFF, (size_t)m * 4, stream));
    r->launches += 1;  // results will be read from this batch
    for (int64_t j = 0; j < m; ++j) bad.push_back((int)j);
  }
  if (int rc = retry_ladder(r.get(), bad)) return rc;
  S.launches += r->launches + r->aux_launches;
  std::vector<double> oState((size_t)m * N), oTau(m);
  std::vector<int64_t> oStats((size_t)m * 8);
  CT_CUDA(cudaMemcpyAsync(oState.data(), r->state.p, (size_t)m * N * 8, cudaMemcpyDeviceToHost, stream));
  CT_CUDA(cudaMemcpyAsync(oTau.data(), r->tau.p, (size_t)m * 8, cudaMemcpyDeviceToHost, stream));
  CT_CUDA(cudaMemcpyAsync(oStats.data(), r->stats.p, (size_t)m * 64, cudaMemcpyDeviceToHost, stream));
  CT_CUDA(cudaMemcpyAsync(oStatus.data(), r->status.p, (size_t)m * 4, cudaMemcpyDeviceToHost, stream));
  CT_CUDA(cudaStreamSynchronize(stream));
  int64_t left = 0;
  for (int64_t j = 0; j < m; ++j) {
    if (oStatus[j] < 0) { ++left; continue; }
    const int64_t i = gidx[j];
    if (state) std::memcpy(state + i * N, oState.data() + j * N, (size_t)N * 8);
    if (tau) tau[i] = oTau[j];
    if (status) status[i] = oStatus[j];
    if (stats) std::memcpy(stats + i * 8, oStats.data() + j * 8, 64);
    if (oStatus[j] != CT_TSTOP_RETURN) S.all_tstop = false;
  }
  S.failed_final = left;
  if (left == 0) S.worst = 0;
  return 0;
}

}  // namespace

extern "C" {

int ct_host_alloc(int64_t bytes, void** out) {
  if (!out || bytes < 0) return fail(CT_ERR_INVALID, "bad argument");
  if (int r = require_device()) return r;
  CT_CUDA(cudaHostAlloc(out, std::max<size_t>((size_t)bytes, 8), cudaHostAllocPortable));
  return 0;
}
int ct_host_free(void* p) {
  if (p) CT_CUDA(cudaFreeHost(p));
  return 0;
}

ct_ensemble* ct_ensemble_create(ct_mech* m, int rt, int64_t n, int n_devices, const int* devices, int* status) {
  auto set = [&](int s) { if (status) *status = s; };
  if (!m || n < 0 || rt < 0 || rt > CT_TABLE_V || n_devices < 0) { set(fail(CT_ERR_INVALID, "bad ensemble arguments")); return nullptr; }
  if (int r = require_device()) { set(r); return nullptr; }
  const int have = ct_device_count();
  std::vector<int> devs;
  if (n_devices == 0) for (int d = 0; d < have; ++d) devs.push_back(d);
  else for (int i = 0; i < n_devices; ++i) devs.push_back(devices ? devices[i] : i);
  for (size_t i = 0; i < devs.size(); ++i) {
    if (devs[i] < 0 || devs[i] >= have) { set(fail(CT_ERR_INVALID, "device ordinal out of range")); return nullptr; }
    for (size_t j = 0; j < i; ++j) if (devs[j] == devs[i]) { set(fail(CT_ERR_INVALID, "a device is listed twice")); return nullptr; }
  }
  ct_ensemble* e = new ct_ensemble;
  e->mech = m; e->rt = rt; e->n = n; e->K = m->C.K; e->N = rt_has_energy(rt) ? m->C.K + 1 : m->C.K;
  e->devices = devs;
  e->shards.resize(devs.size());
  set(0);
  return e;
}
void ct_ensemble_destroy(ct_ensemble* e) { delete e; }
int ct_ensemble_n_shards(const ct_ensemble* e) { return e ? (int)e->devices.size() : fail(CT_ERR_INVALID, "null ensemble"); }
int ct_ensemble_set_partition(ct_ensemble* e, int mode, int64_t chunk) {
  if (!e) return fail(CT_ERR_INVALID, "null ensemble");
  if ((mode != CT_PARTITION_BLOCK && mode != CT_PARTITION_CYCLIC) || chunk < 0) return fail(CT_ERR_INVALID, "bad partition arguments");
  e->partition = mode;
  if (chunk > 0) e->chunk = chunk;
  return 0;
}
int ct_ensemble_set_tolerances(ct_ensemble* e, double rtol, double atol) {
  if (!e) return fail(CT_ERR_INVALID, "null ensemble");
  if (!(rtol >= 0.0) || !(atol >= 0.0) || (rtol == 0.0 && atol == 0.0)) return fail(CT_ILL_INPUT, "tolerances must be non-negative and not both zero");
  e->rtol = rtol; e->atol = atol;
  return 0;
}
int ct_ensemble_set_stop_time(ct_ensemble* e, double t_end) {
  if (!e) return fail(CT_ERR_INVALID, "null ensemble");
  if (!(t_end > 0.0)) return fail(CT_ILL_INPUT, "stop time must be positive");
  e->t_stop = t_end;
  return 0;
}
int ct_ensemble_set_max_steps(ct_ensemble* e, int64_t mx) {
  if (!e) return fail(CT_ERR_INVALID, "null ensemble");
  if (mx < 0) return fail(CT_ILL_INPUT, "negative maximum number of steps");
  e->mxstep = mx == 0 ? 500 : mx;
  return 0;
}
int ct_ensemble_set_first_pass_steps(ct_ensemble* e, int64_t steps) {
  if (!e) return fail(CT_ERR_INVALID, "null ensemble");
  if (steps < 0) return fail(CT_ILL_INPUT, "negative number of steps");
  e->first_pass = steps;
  return 0;
}
int ct_ensemble_set_retry(ct_ensemble* e, int on) { if (!e) return fail(CT_ERR_INVALID, "null ensemble"); e->retry = on ? 1 : 0; return 0; }
int ct_ensemble_set_streams(ct_ensemble* e, int streams) {
  if (!e) return fail(CT_ERR_INVALID, "null ensemble");
  if (streams < 1 || streams > 4) return fail(CT_ERR_INVALID, "1 to 4 launches in flight per device");
  e->streams = streams;
  return 0;
}
int ct_ensemble_set_hot_first(ct_ensemble* e, int on) { if (!e) return fail(CT_ERR_INVALID, "null ensemble"); e->hot_first = on ? 1 : 0; return 0; }
int ct_ensemble_set_configure(ct_ensemble* e, ct_batch_configure_fn fn, void* user) {
  if (!e) return fail(CT_ERR_INVALID, "null ensemble");
  e->configure = fn; e->configure_user = user;
  return 0;
}
int ct_ensemble_set_tp_table(ct_ensemble* e, int nt, const double* t, const double* T, const double* p) {
  if (!e) return fail(CT_ERR_INVALID, "null ensemble");
  if (!rt_has_table(e->rt)) return fail(CT_ILL_INPUT, "tables belong to CT_TABLE_TP (T, p) and CT_TABLE_V (V, dV/dt) reactors");
  if (nt < 2 || !t || !T || !p) return fail(CT_ERR_INVALID, "bad table arguments");
  e->nt = nt; e->tab_t = t; e->tab_T = T; e->tab_p = p;
  return 0;
}
int ct_partition_range(int64_t n, int n_shards, int mode, int shard, int64_t* first, int64_t* stride, int64_t* count) {
  if (n < 0 || n_shards < 1 || shard < 0 || shard >= n_shards || (mode != CT_PARTITION_BLOCK && mode != CT_PARTITION_CYCLIC))
    return fail(CT_ERR_INVALID, "bad partition arguments");
  ct_ensemble e;
  e.n = n; e.partition = mode; e.devices.assign(n_shards, 0);
  if (first) *first = e.global_index(shard, 0);
  if (stride) *stride = mode == CT_PARTITION_CYCLIC ? (int64_t)n_shards : 1;
  if (count) *count = e.shard_count(shard);
  return 0;
}
int ct_ensemble_shard_range(const ct_ensemble* e, int shard, int64_t* first, int64_t* stride, int64_t* count) {
  if (!e) return fail(CT_ERR_INVALID, "null ensemble");
  return ct_partition_range(e->n, (int)e->devices.size(), e->partition, shard, first, stride, count);
}

int ct_ensemble_integrate(ct_ensemble* e, const double* T, const double* p, const double* Y, double* state, double* tau,
                          int32_t* status, int64_t* stats) {
  if (!e || !T || !p || !Y) return fail(CT_ERR_INVALID, "null argument");
  if (rt_has_table(e->rt) && e->nt < 2) return fail(CT_ILL_INPUT, "table ensemble (CT_TABLE_TP / CT_TABLE_V) without its table");
  const int G = (int)e->devices.size();
  if (G == 0) return fail(CT_ERR_NO_DEVICE, "no device in the ensemble");
  int prev = 0;
  CT_CUDA(cudaGetDevice(&prev));
  for (int g = 0; g < G; ++g) { e->shards[g] = ct_ensemble::Shard(); e->shards[g].count = e->shard_count(g); }
  const auto t_begin = std::chrono::steady_clock::now();
  {
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
      const int64_t nchunks = (e->shards[g].count + e->chunk - 1) / std::max<int64_t>(e->chunk, 1);
      const int W = (int)std::max<int64_t>(1, std::min<int64_t>(e->streams, nchunks));
      for (int w = 0; w < W; ++w) {
        th.emplace_back([=]() {
          ct_ensemble::Shard& S = e->shards[g];
          const int rc = S.count > 0 ? run_worker(e, g, w, W, T, p, Y, state, tau, status, stats) : 0;
          if (rc < 0) { std::lock_guard<std::mutex> lock(S.mu); if (S.rc == 0) { S.rc = rc; S.err = g_err; } }  // thread-local message of the worker
        });
      }
    }
    for (auto& t : th) t.join();
  }
  {
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
      ct_ensemble::Shard& S = e->shards[g];
      if (S.rc < 0) continue;
      if (!e->retry || S.bad.empty()) { S.failed_final = S.failed_first; continue; }
      th.emplace_back([=]() {
        ct_ensemble::Shard& Sg = e->shards[g];
        const int rc = run_ladder(e, g, T, p, Y, state, tau, status, stats);
        if (rc < 0) { Sg.rc = rc; Sg.err = g_err; }
      });
    }
    for (auto& t : th) t.join();
  }
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
  for (int g = 0; g < G; ++g) e->shards[g].wall_s = wall;
  cudaSetDevice(prev);
  int worst = 0;
  bool all_tstop = true, any = false;
  for (int g = 0; g < G; ++g) {
    const ct_ensemble::Shard& S = e->shards[g];
    if (S.rc < 0) return fail(S.rc, "shard " + std::to_string(g) + " (device " + std::to_string(e->devices[g]) + "): " + S.err);
    if (S.count > 0) { any = true; worst = std::min(worst, S.worst); all_tstop = all_tstop && S.all_tstop; }
  }
  return worst < 0 ? worst : (any && all_tstop ? CT_TSTOP_RETURN : CT_SUCCESS);
}

int ct_ensemble_shard_info(const ct_ensemble* e, int shard, int64_t counts[5], double times[2]) {
  if (!e || shard < 0 || shard >= (int)e->devices.size()) return fail(CT_ERR_INVALID, "bad shard");
  const ct_ensemble::Shard& S = e->shards[shard];
  if (counts) { counts[0] = S.count; counts[1] = S.chunks; counts[2] = S.launches; counts[3] = S.failed_first; counts[4] = S.failed_final; }
  if (times) { times[0] = S.kernel_ms; times[1] = S.wall_s; }
  return 0;
}

}  // extern "C"
