/* ORACLE — test infrastructure, NOT product code (see ct_oracle.h).
 *
 * Batch driver: what a host program written against the reference would do
 * for an ensemble — one Apps::Reactor + one CVODE instance per case,
 * integrated one after the other (tests/src/reactor.cpp:61-126 of the
 * reference, repeated per case) — spread over host threads, one reactor per
 * core.  Used as bench.py's cpu_baseline and `--impl reference` leg.
 */
#include "ct_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>

static int rhs_tramp(double t, const double* y, double* ydot, void* user) {
  return or_rhs((oreactor*)user, t, y, ydot);
}

typedef struct {
  const omech* m; int type; long n; const double *T0, *p0, *Y0; int nt; const double *tab_t, *tab_T, *tab_p;
  double rtol, atol, t_end; long max_steps; int ign_mode; double ign_value; int n_out;
  double *out_traj, *out_state, *out_tau; long* out_stats; int* out_status;
  const double* opts;
  double* out_stats_ex; /* optional [n x 6]: qlast, qcur, hinused, hlast, hcur, tcur (MainSolverStatistics of the reference) */
  atomic_long next;
} batch_job;

/* set by or_set_stats_ex_buffer for the next or_integrate_batch* call of this thread (keeps the call signature) */
static _Thread_local double* g_stats_ex = NULL;
void or_set_stats_ex_buffer(double* buf) { g_stats_ex = buf; }

static void run_one(batch_job* J, long i) {
  const omech* m = J->m;
  const int K = om_nspecies(m);
  oreactor* r = or_create(m, J->type, J->T0[i], J->p0[i], J->Y0 + i * K);
  if (J->type == OR_TABLE_TP || J->type == OR_TABLE_V) or_set_table(r, J->nt, J->tab_t, J->tab_T + i * J->nt, J->tab_p + i * J->nt);
  const int N = or_nstate(r);
  const int n_out = J->n_out;
  double* y = (double*)malloc(N * 8);
  or_initial_state(r, y);
  cvo* c = cvo_create(N, rhs_tramp, NULL, r);
  cvo_init(c, 0.0, y);
  cvo_set_tolerances(c, J->rtol, J->atol, NULL);
  cvo_set_stop_time(c, J->t_end);
  cvo_set_max_steps(c, J->max_steps);
  if (J->opts) {
    const double* o = J->opts;
    cvo_set_step_options(c, o[0], o[1], o[2]);
    cvo_set_solver_options(c, (int)o[4], (int)o[5], (int)o[6], o[3]);
    if (o[7] > 0) cvo_set_max_order(c, (int)o[7]);
  }
  if (J->type <= OR_CONST_VOL || J->type == OR_TABLE_V) cvo_watch(c, 0, J->ign_mode == 0 ? J->ign_value : J->T0[i] + J->ign_value);
  int status = 0;
  double tret = 0.0;
  if (n_out > 0) {
    /* output cadence of tests/src/reactor.cpp:105-126: t_k = (k+1)*t_end/n_out */
    for (int k = 0; k < n_out; ++k) {
      double tout = (k + 1 == n_out) ? J->t_end : (k + 1) * (J->t_end / n_out);
      status = cvo_integrate(c, tout, y, &tret);
      memcpy(J->out_traj + ((size_t)i * n_out + k) * N, y, N * 8);
      if (status < 0) {
        for (int kk = k + 1; kk < n_out; ++kk) memcpy(J->out_traj + ((size_t)i * n_out + kk) * N, y, N * 8);
        break;
      }
    }
  } else {
    status = cvo_integrate(c, J->t_end, y, &tret);
  }
  memcpy(J->out_state + (size_t)i * N, y, N * 8);
  if (J->out_tau) J->out_tau[i] = cvo_watch_time(c);
  if (J->out_stats) {
    long io[10]; double ro[4];
    cvo_stats(c, io, ro);
    for (int s = 0; s < 8; ++s) J->out_stats[i * 8 + s] = io[s];
    if (J->out_stats_ex) {
      double* e = J->out_stats_ex + i * 6;
      e[0] = (double)io[8]; e[1] = (double)io[9]; e[2] = ro[2]; e[3] = ro[0]; e[4] = ro[1]; e[5] = ro[3];
    }
  }
  if (J->out_status) J->out_status[i] = status;
  cvo_destroy(c);
  or_destroy(r);
  free(y);
}

static void* worker(void* arg) {
  batch_job* J = (batch_job*)arg;
  for (;;) {
    long i = atomic_fetch_add(&J->next, 1);
    if (i >= J->n) break;
    run_one(J, i);
  }
  return NULL;
}

int or_integrate_batch(const omech* m, int type, long n, const double* T0, const double* p0,
                       const double* Y0, int nt, const double* tab_t, const double* tab_T,
                       const double* tab_p, double rtol, double atol, double t_end, long max_steps,
                       int ign_mode, double ign_value, int nthreads, int n_out, double* out_traj,
                       double* out_state, double* out_tau, long* out_stats, int* out_status) {
  return or_integrate_batch_opts(m, type, n, T0, p0, Y0, nt, tab_t, tab_T, tab_p, rtol, atol, t_end, max_steps, ign_mode,
                                 ign_value, nthreads, n_out, out_traj, out_state, out_tau, out_stats, out_status, NULL);
}

int or_integrate_batch_opts(const omech* m, int type, long n, const double* T0, const double* p0,
                            const double* Y0, int nt, const double* tab_t, const double* tab_T,
                            const double* tab_p, double rtol, double atol, double t_end, long max_steps,
                            int ign_mode, double ign_value, int nthreads, int n_out, double* out_traj,
                            double* out_state, double* out_tau, long* out_stats, int* out_status,
                            const double* opts) {
  batch_job J = {m, type, n, T0, p0, Y0, nt, tab_t, tab_T, tab_p, rtol, atol, t_end, max_steps,
                 ign_mode, ign_value, n_out, out_traj, out_state, out_tau, out_stats, out_status, opts, g_stats_ex, 0};
  g_stats_ex = NULL;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = (int)(n > 0 ? n : 1);
  if (nthreads == 1) { worker(&J); return 0; }
  pthread_t* th = (pthread_t*)malloc(nthreads * sizeof(pthread_t));
  for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, &J);
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th);
  return 0;
}
