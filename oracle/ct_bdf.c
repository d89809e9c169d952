/* ORACLE — test infrastructure, NOT product code (see ct_oracle.h).
 *
 * CPU restatement of the CVODE algorithm the reference drives through
 * SUNDIALS::CVODE::IStrategy::Initialise / Integrate
 * (/root/reference/src/sundials/cvode/interface.cpp:222-427, :454-475) with the
 * Dense strategy (:714-721) and no user Jacobian (:698-703 -> CVODE's internal
 * difference-quotient dense Jacobian): variable-order (1..5) variable-step
 * fixed-leading-coefficient BDF in Nordsieck form, modified Newton with a saved
 * LU of I - gamma*J, WRMS error control, CV_NORMAL return with tstop.
 * SUNDIALS (pinned v5.8.0) is not in /root/reference; this follows its
 * published algorithm (Hindmarsh et al., CVODE v5 user guide / cvode.c structure:
 * cvHin, cvStep, cvPredict, cvSetBDF, cvNls + Newton, cvDoErrorTest,
 * cvCompleteStep, cvPrepareNextStep, cvLsSetup, cvLsDenseDQJac, CVodeGetDky)
 * with its default constants (types.hpp:82-111 of the reference lists the same
 * defaults: max order 5, 3 Newton iterations, coefficient 0.1, 7 error-test
 * failures, 10 convergence failures, 500 steps).
 * Pinned by tests/test_oracle_roberts.py against the reference's golden file
 * tests/data/sundials_cvode/direct_dense_cvRoberts_dns.dat.
 */
#include "ct_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define QMAX 5
#define LMAX (QMAX + 1)
#define UROUND DBL_EPSILON

/* CVODE constants */
#define FUZZ_FACTOR 100.0
#define HLB_FACTOR 100.0
#define HUB_FACTOR 0.1
#define H_BIAS 0.5
#define HIN_MAX_ITERS 4
#define ETAMX1 10000.0
#define ETAMX2 10.0
#define ETAMX3 10.0
#define ETAMXF 0.2
#define ETAMIN 0.1
#define ETACF 0.25
#define ADDON 0.000001
#define BIAS1 6.0
#define BIAS2 6.0
#define BIAS3 10.0
#define THRESH 1.5
#define ONEPSM 1.000001
#define SMALL_NST 10
#define MXNEF1 3
#define SMALL_NEF 2
#define LONG_WAIT 10
#define CRDOWN 0.3
#define DGMAX 0.3
#define RDIV 2.0
#define MSBP 20
#define MSBJ 50
#define LS_DGMAX 0.2
#define MIN_INC_MULT 1000.0

enum { FIRST_CALL = 101, PREV_CONV_FAIL = 102, PREV_ERR_FAIL = 103 };
enum { DO_ERROR_TEST = 2, PREDICT_AGAIN = 3, TRY_AGAIN = 5, CONV_RECVR = 901 };
enum { NO_FAILURES = 0, FAIL_BAD_J = 1, FAIL_OTHER = 2 };

struct cvo {
  int N;
  cvo_rhs_fn f;
  cvo_jac_fn jac;
  void* user;
  double rtol, atol_s;
  double* atol_v;
  double* zn[LMAX + 1]; /* zn[0..qmax] */
  double *ewt, *y, *acor, *tempv, *ftemp, *vtemp;
  double *M, *savedJ;
  int* piv;
  int q, qprime, qwait, L, qmax, qu;
  double h, hprime, eta, hscale, tn, hu, h0u, next_h;
  int next_q; /* CVodeGetCurrentOrder: order to be attempted on the next step (cv_next_q) */
  double tau[LMAX + 1], tq[6], l[LMAX];
  double rl1, gamma, gammap, gamrat, crate, delp, acnrm;
  int acnrmcur;
  double etamax, etaq, etaqm1, etaqp1, saved_tq5;
  double hmin, hmax_inv, nlscoef, hin; /* hin: user initial step (CVodeSetInitStep), 0 = estimate */
  int maxnef, maxncf, maxcor;
  long mxstep, nst, nfe, nsetups, netf, nni, ncfn, nje, nfeDQ, nstlp, nstlj;
  int tstopset; double tstop;
  int jcur, convfail;
  int watch_comp; double watch_thr, watch_time; int watch_done;
};

static double wrms(int N, const double* v, const double* w) {
  double s = 0.0;
  for (int i = 0; i < N; ++i) { double p = v[i] * w[i]; s += p * p; }
  return sqrt(s / N);
}

static int ewt_set(cvo* c, const double* y) {
  for (int i = 0; i < c->N; ++i) {
    double a = c->atol_v ? c->atol_v[i] : c->atol_s;
    double t = c->rtol * fabs(y[i]) + a;
    if (t <= 0.0) return -1;
    c->ewt[i] = 1.0 / t;
  }
  return 0;
}

cvo* cvo_create(int N, cvo_rhs_fn f, cvo_jac_fn jac, void* user) {
  cvo* c = (cvo*)calloc(1, sizeof(cvo));
  c->N = N; c->f = f; c->jac = jac; c->user = user;
  for (int j = 0; j <= LMAX; ++j) c->zn[j] = (double*)calloc(N, 8);
  c->ewt = (double*)calloc(N, 8); c->y = (double*)calloc(N, 8); c->acor = (double*)calloc(N, 8);
  c->tempv = (double*)calloc(N, 8); c->ftemp = (double*)calloc(N, 8); c->vtemp = (double*)calloc(N, 8);
  c->M = (double*)calloc((size_t)N * N, 8); c->savedJ = (double*)calloc((size_t)N * N, 8);
  c->piv = (int*)calloc(N, sizeof(int));
  c->qmax = QMAX; c->mxstep = 500; c->maxnef = 7; c->maxncf = 10; c->maxcor = 3; c->nlscoef = 0.1;
  c->rtol = 1e-4; c->atol_s = 1e-8;
  c->watch_comp = -1; c->watch_time = NAN;
  return c;
}
void cvo_destroy(cvo* c) {
  if (!c) return;
  for (int j = 0; j <= LMAX; ++j) free(c->zn[j]);
  free(c->ewt); free(c->y); free(c->acor); free(c->tempv); free(c->ftemp); free(c->vtemp);
  free(c->M); free(c->savedJ); free(c->piv); free(c->atol_v); free(c);
}
void cvo_init(cvo* c, double t0, const double* y0) {
  memcpy(c->zn[0], y0, c->N * 8);
  c->tn = t0; c->q = 1; c->L = 2; c->qwait = 2; c->etamax = ETAMX1; c->qu = 0; c->hu = 0.0;
  c->nst = c->nfe = c->nsetups = c->netf = c->nni = c->ncfn = c->nje = c->nfeDQ = 0;
  c->nstlp = c->nstlj = 0; c->h = 0.0; c->h0u = 0.0; c->next_h = 0.0; c->next_q = 0; c->tstopset = 0;
  c->crate = 1.0; c->saved_tq5 = 0.0; c->gamrat = 1.0; c->gammap = 0.0;
  c->watch_done = 0; c->watch_time = NAN;
  for (int i = 0; i <= LMAX; ++i) c->tau[i] = 0.0;
}
void cvo_set_tolerances(cvo* c, double rtol, double atol_s, const double* atol_v) {
  c->rtol = rtol; c->atol_s = atol_s;
  free(c->atol_v); c->atol_v = NULL;
  if (atol_v) { c->atol_v = (double*)malloc(c->N * 8); memcpy(c->atol_v, atol_v, c->N * 8); }
}
void cvo_set_stop_time(cvo* c, double tstop) { c->tstop = tstop; c->tstopset = 1; }
void cvo_set_max_steps(cvo* c, long mx) { c->mxstep = mx > 0 ? mx : 500; }
void cvo_set_max_order(cvo* c, int q) { if (q >= 1 && q <= QMAX) c->qmax = q; }
/* CVodeSetInitStep / SetMinStep / SetMaxStep (IStrategy::SetInitStepSize..., interface.cpp:61-97 of the reference):
 * 0 keeps CVODE's default (estimated h0, hmin = 0, hmax = inf) */
void cvo_set_step_options(cvo* c, double h_init, double h_min, double h_max) {
  c->hin = h_init; c->hmin = h_min > 0.0 ? h_min : 0.0; c->hmax_inv = h_max > 0.0 ? 1.0 / h_max : 0.0;
}
/* CVodeSetMaxErrTestFails / SetMaxNonlinIters / SetMaxConvFails / SetNonlinConvCoef (interface.cpp:166-219);
 * non-positive values keep the defaults 7 / 3 / 10 / 0.1 */
void cvo_set_solver_options(cvo* c, int maxnef, int maxcor, int maxncf, double nlscoef) {
  if (maxnef > 0) c->maxnef = maxnef;
  if (maxcor > 0) c->maxcor = maxcor;
  if (maxncf > 0) c->maxncf = maxncf;
  if (nlscoef > 0.0) c->nlscoef = nlscoef;
}
void cvo_watch(cvo* c, int comp, double thr) { c->watch_comp = comp; c->watch_thr = thr; c->watch_done = 0; c->watch_time = NAN; }
double cvo_watch_time(const cvo* c) { return c->watch_time; }
void cvo_stats(const cvo* c, long* io, double* ro) {
  io[0] = c->nst; io[1] = c->nfe; io[2] = c->nsetups; io[3] = c->netf; io[4] = c->nni;
  io[5] = c->ncfn; io[6] = c->nje; io[7] = c->nfeDQ; io[8] = c->qu; io[9] = c->next_q;
  ro[0] = c->hu; ro[1] = c->next_h; ro[2] = c->h0u; ro[3] = c->tn;
}

/* ---- dense LU with partial pivoting (SUNLinSol_Dense: denseGETRF/GETRS), column-major ---- */
static int dense_getrf(double* a, int n, int* p) {
  for (int k = 0; k < n; ++k) {
    double* col_k = a + (size_t)k * n;
    int l = k;
    for (int i = k + 1; i < n; ++i) if (fabs(col_k[i]) > fabs(col_k[l])) l = i;
    p[k] = l;
    if (col_k[l] == 0.0) return k + 1;
    if (l != k)
      for (int j = 0; j < n; ++j) { double t = a[(size_t)j * n + l]; a[(size_t)j * n + l] = a[(size_t)j * n + k]; a[(size_t)j * n + k] = t; }
    double mult = 1.0 / col_k[k];
    for (int i = k + 1; i < n; ++i) col_k[i] *= mult;
    for (int j = k + 1; j < n; ++j) {
      double* col_j = a + (size_t)j * n;
      double a_kj = col_j[k];
      if (a_kj != 0.0) for (int i = k + 1; i < n; ++i) col_j[i] -= a_kj * col_k[i];
    }
  }
  return 0;
}
static void dense_getrs(const double* a, int n, const int* p, double* b) {
  for (int k = 0; k < n; ++k) { int pk = p[k]; if (pk != k) { double t = b[k]; b[k] = b[pk]; b[pk] = t; } }
  for (int k = 0; k < n - 1; ++k) { const double* col_k = a + (size_t)k * n; for (int i = k + 1; i < n; ++i) b[i] -= col_k[i] * b[k]; }
  for (int k = n - 1; k > 0; --k) { const double* col_k = a + (size_t)k * n; b[k] /= col_k[k]; for (int i = 0; i < k; ++i) b[i] -= col_k[i] * b[k]; }
  b[0] /= a[0];
}

/* ---- cvHin and helpers ---- */
static double upper_bound_h0(cvo* c, double tdist) {
  int N = c->N; double hub_inv = 0.0;
  double* t1 = c->tempv;
  ewt_set(c, c->zn[0]);
  for (int i = 0; i < N; ++i) {
    double v = HUB_FACTOR * fabs(c->zn[0][i]) + 1.0 / c->ewt[i];
    (void)t1;
    double r = fabs(c->zn[1][i]) / v;
    if (r > hub_inv) hub_inv = r;
  }
  double hub = HUB_FACTOR * tdist;
  if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
  return hub;
}
static int ydd_norm(cvo* c, double hg, double* yddnrm) {
  int N = c->N;
  for (int i = 0; i < N; ++i) c->y[i] = hg * c->zn[1][i] + c->zn[0][i];
  int r = c->f(c->tn + hg, c->y, c->tempv, c->user); c->nfe++;
  if (r) return r;
  for (int i = 0; i < N; ++i) c->tempv[i] = (1.0 / hg) * c->tempv[i] - (1.0 / hg) * c->zn[1][i];
  *yddnrm = wrms(N, c->tempv, c->ewt);
  return 0;
}
static int cv_hin(cvo* c, double tout) {
  double tdiff = tout - c->tn;
  if (tdiff == 0.0) return CVO_TOO_CLOSE;
  int sign = tdiff > 0.0 ? 1 : -1;
  double tdist = fabs(tdiff);
  double tround = UROUND * fmax(fabs(c->tn), fabs(tout));
  if (tdist < 2.0 * tround) return CVO_TOO_CLOSE;
  double hlb = HLB_FACTOR * tround;
  double hub = upper_bound_h0(c, tdist);
  double hg = sqrt(hlb * hub);
  if (hub < hlb) { c->h = sign == -1 ? -hg : hg; return 0; }
  int hnewOK = 0; double hs = hg, hnew = hg, yddnrm = 0.0;
  for (int count1 = 1; count1 <= HIN_MAX_ITERS; ++count1) {
    int hgOK = 0;
    for (int count2 = 1; count2 <= HIN_MAX_ITERS; ++count2) {
      double hgs = hg * sign;
      int r = ydd_norm(c, hgs, &yddnrm);
      if (r < 0) return -8;
      if (r == 0) { hgOK = 1; break; }
      hg *= 0.2;
    }
    if (!hgOK) { if (count1 <= 2) return -10; hnew = hs; break; }
    hs = hg;
    if (hnewOK || count1 == HIN_MAX_ITERS) { hnew = hg; break; }
    hnew = (yddnrm * hub * hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(hg * hub);
    double hrat = hnew / hg;
    if (hrat > 0.5 && hrat < 2.0) hnewOK = 1;
    if (count1 > 1 && hrat > 2.0) { hnew = hg; hnewOK = 1; }
    hg = hnew;
  }
  double h0 = H_BIAS * hnew;
  if (h0 < hlb) h0 = hlb;
  if (h0 > hub) h0 = hub;
  if (sign == -1) h0 = -h0;
  c->h = h0;
  return 0;
}

/* ---- order/step adjustment ---- */
static void cv_rescale(cvo* c) {
  double factor = c->eta;
  for (int j = 1; j <= c->q; ++j) { for (int i = 0; i < c->N; ++i) c->zn[j][i] *= factor; factor *= c->eta; }
  c->h = c->hscale * c->eta; c->next_h = c->h; c->hscale = c->h;
}
static void cv_increase_bdf(cvo* c) {
  double l[LMAX + 1];
  for (int i = 0; i <= c->qmax; ++i) l[i] = 0.0;
  double alpha1, prod, xiold, alpha0, hsum, xi, A1;
  l[2] = alpha1 = prod = xiold = 1.0; alpha0 = -1.0; hsum = c->hscale;
  if (c->q > 1) {
    for (int j = 1; j < c->q; ++j) {
      hsum += c->tau[j + 1]; xi = hsum / c->hscale; prod *= xi;
      alpha0 -= 1.0 / (j + 1); alpha1 += 1.0 / xi;
      for (int i = j + 2; i >= 2; --i) l[i] = l[i] * xiold + l[i - 1];
      xiold = xi;
    }
  }
  A1 = (-alpha0 - alpha1) / prod;
  int L = c->L;
  for (int i = 0; i < c->N; ++i) c->zn[L][i] = A1 * c->zn[c->qmax][i];
  for (int j = 2; j <= c->q; ++j) for (int i = 0; i < c->N; ++i) c->zn[j][i] += l[j] * c->zn[L][i];
}
static void cv_decrease_bdf(cvo* c) {
  double l[LMAX + 1];
  for (int i = 0; i <= c->qmax; ++i) l[i] = 0.0;
  l[2] = 1.0; double hsum = 0.0, xi;
  for (int j = 1; j <= c->q - 2; ++j) {
    hsum += c->tau[j]; xi = hsum / c->hscale;
    for (int i = j + 2; i >= 2; --i) l[i] = l[i] * xi + l[i - 1];
  }
  for (int j = 2; j < c->q; ++j) for (int i = 0; i < c->N; ++i) c->zn[j][i] += -l[j] * c->zn[c->q][i];
}
static void cv_adjust_order(cvo* c, int deltaq) {
  if (c->q == 2 && deltaq != 1) return;
  if (deltaq == 1) cv_increase_bdf(c); else if (deltaq == -1) cv_decrease_bdf(c);
}
static void cv_adjust_params(cvo* c) {
  if (c->qprime != c->q) { cv_adjust_order(c, c->qprime - c->q); c->q = c->qprime; c->L = c->q + 1; c->qwait = c->L; }
  cv_rescale(c);
}
static void cv_predict(cvo* c) {
  c->tn += c->h;
  if (c->tstopset && (c->tn - c->tstop) * c->h > 0.0) c->tn = c->tstop;
  for (int k = 1; k <= c->q; ++k) for (int j = c->q; j >= k; --j) for (int i = 0; i < c->N; ++i) c->zn[j - 1][i] += c->zn[j][i];
}
static void cv_restore(cvo* c, double saved_t) {
  c->tn = saved_t;
  for (int k = 1; k <= c->q; ++k) for (int j = c->q; j >= k; --j) for (int i = 0; i < c->N; ++i) c->zn[j - 1][i] -= c->zn[j][i];
}
static void cv_set_bdf(cvo* c) {
  int q = c->q; double* l = c->l; double h = c->h;
  double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
  l[0] = l[1] = xi_inv = xistar_inv = 1.0;
  for (int i = 2; i <= q; ++i) l[i] = 0.0;
  alpha0 = alpha0_hat = -1.0; hsum = h;
  if (q > 1) {
    for (int j = 2; j < q; ++j) {
      hsum += c->tau[j - 1]; xi_inv = h / hsum; alpha0 -= 1.0 / j;
      for (int i = j; i >= 1; --i) l[i] += l[i - 1] * xi_inv;
    }
    alpha0 -= 1.0 / q; xistar_inv = -l[1] - alpha0; hsum += c->tau[q - 1]; xi_inv = h / hsum;
    alpha0_hat = -l[1] - xi_inv;
    for (int i = q; i >= 1; --i) l[i] += l[i - 1] * xistar_inv;
  }
  /* cvSetTqBDF */
  double A1 = 1.0 - alpha0_hat + alpha0, A2 = 1.0 + q * A1;
  c->tq[2] = fabs(A1 / (alpha0 * A2));
  c->tq[5] = fabs(A2 * xistar_inv / (l[q] * xi_inv));
  if (c->qwait == 1) {
    if (q > 1) {
      double C = xistar_inv / l[q], A3 = alpha0 + 1.0 / q, A4 = alpha0_hat + xi_inv;
      double Cpinv = (1.0 - A4 + A3) / A3;
      c->tq[1] = fabs(C * Cpinv);
    } else c->tq[1] = 1.0;
    hsum += c->tau[q]; xi_inv = h / hsum;
    double A5 = alpha0 - (1.0 / (q + 1)), A6 = alpha0_hat - xi_inv;
    double Cppinv = (1.0 - A6 + A5) / A2;
    c->tq[3] = fabs(Cppinv / (xi_inv * (q + 2) * A5));
  }
  c->tq[4] = c->nlscoef / c->tq[2];
  /* tail of cvSet */
  c->rl1 = 1.0 / l[1]; c->gamma = h * c->rl1;
  if (c->nst == 0) c->gammap = c->gamma;
  c->gamrat = (c->nst > 0) ? c->gamma / c->gammap : 1.0;
}

/* ---- linear solver interface (cvLsSetup/cvLsLinSys/cvLsDenseDQJac/cvLsSolve) ---- */
static int dq_jac(cvo* c, double t, double* y, const double* fy, double* J) {
  int N = c->N;
  double srur = sqrt(UROUND);
  double fnorm = wrms(N, fy, c->ewt);
  double minInc = (fnorm != 0.0) ? (MIN_INC_MULT * fabs(c->h) * UROUND * N * fnorm) : 1.0;
  for (int j = 0; j < N; ++j) {
    double yj = y[j];
    double inc = fmax(srur * fabs(yj), minInc / c->ewt[j]);
    y[j] += inc;
    int r = c->f(t, y, c->vtemp, c->user); c->nfeDQ++;
    y[j] = yj;
    if (r) return r;
    double inc_inv = 1.0 / inc;
    for (int i = 0; i < N; ++i) J[(size_t)j * N + i] = inc_inv * c->vtemp[i] - inc_inv * fy[i];
  }
  return 0;
}
static int ls_setup(cvo* c, int convfail) {
  int N = c->N; size_t nn = (size_t)N * N;
  double dgamma = fabs((c->gamma / c->gammap) - 1.0);
  int jbad = (c->nst == 0) || (c->nst > c->nstlj + MSBJ) || ((convfail == FAIL_BAD_J) && (dgamma < LS_DGMAX)) || (convfail == FAIL_OTHER);
  if (!jbad) {
    c->jcur = 0;
    memcpy(c->M, c->savedJ, nn * 8);
  } else {
    c->jcur = 1;
    memset(c->M, 0, nn * 8);
    int r = c->jac ? c->jac(c->tn, c->y, c->ftemp, c->M, c->user) : dq_jac(c, c->tn, c->y, c->ftemp, c->M);
    c->nje++; c->nstlj = c->nst;
    if (r < 0) return -1;
    if (r > 0) return 1;
    memcpy(c->savedJ, c->M, nn * 8);
  }
  for (size_t i = 0; i < nn; ++i) c->M[i] *= -c->gamma;
  for (int i = 0; i < N; ++i) c->M[(size_t)i * N + i] += 1.0;
  int ier = dense_getrf(c->M, N, c->piv);
  return ier ? 1 : 0;
}

/* ---- Newton (SUNNonlinSol_Newton + cvNlsResidual/LSetup/LSolve/ConvTest) ---- */
static int nls_residual(cvo* c, double* delta) {
  int N = c->N;
  for (int i = 0; i < N; ++i) c->y[i] = c->zn[0][i] + c->acor[i];
  int r = c->f(c->tn, c->y, c->ftemp, c->user); c->nfe++;
  if (r < 0) return -8;
  if (r > 0) return CONV_RECVR;
  for (int i = 0; i < N; ++i) delta[i] = (c->rl1 * c->zn[1][i] + c->acor[i]) - c->gamma * c->ftemp[i];
  return 0;
}
static int cv_nls(cvo* c, int nflag) {
  int N = c->N;
  c->convfail = (nflag == FIRST_CALL || nflag == PREV_ERR_FAIL) ? NO_FAILURES : FAIL_OTHER;
  int callSetup = (nflag == PREV_CONV_FAIL) || (nflag == PREV_ERR_FAIL) || (c->nst == 0) ||
                  (c->nst >= c->nstlp + MSBP) || (fabs(c->gamrat - 1.0) > DGMAX);
  for (int i = 0; i < N; ++i) c->acor[i] = 0.0;
  double* delta = c->tempv;
  int jbad = 0, retval;
  c->acnrmcur = 0;
  for (;;) {
    retval = nls_residual(c, delta);
    if (retval != 0) break;
    if (callSetup) {
      if (jbad) c->convfail = FAIL_BAD_J;
      int r = ls_setup(c, c->convfail);
      c->nsetups++;
      callSetup = 0; /* forceSetup reset */
      c->gamrat = 1.0; c->gammap = c->gamma; c->crate = 1.0; c->nstlp = c->nst;
      if (r < 0) { retval = CVO_LSETUP_FAIL; break; }
      if (r > 0) { retval = CONV_RECVR; break; }
    }
    int m = 0;
    for (;;) {
      c->nni++;
      for (int i = 0; i < N; ++i) delta[i] = -delta[i];
      dense_getrs(c->M, N, c->piv, delta);
      if (c->gamrat != 1.0) { double s = 2.0 / (1.0 + c->gamrat); for (int i = 0; i < N; ++i) delta[i] *= s; }
      for (int i = 0; i < N; ++i) c->acor[i] += delta[i];
      /* cvNlsConvTest */
      double del = wrms(N, delta, c->ewt);
      if (m > 0) c->crate = fmax(CRDOWN * c->crate, del / c->delp);
      double dcon = del * fmin(1.0, c->crate) / c->tq[4];
      if (dcon <= 1.0) {
        c->acnrm = (m == 0) ? del : wrms(N, c->acor, c->ewt);
        c->acnrmcur = 1;
        retval = 0; goto converged;
      }
      if (m >= 1 && del > RDIV * c->delp) { retval = CONV_RECVR; break; }
      c->delp = del;
      m++;
      if (m >= c->maxcor) { retval = CONV_RECVR; break; }
      retval = nls_residual(c, delta);
      if (retval != 0) break;
    }
    if (retval > 0 && !c->jcur) {
      callSetup = 1; jbad = 1;
      for (int i = 0; i < N; ++i) c->acor[i] = 0.0;
      continue;
    }
    break;
  }
  return retval;
converged:
  c->jcur = 0;
  for (int i = 0; i < N; ++i) c->y[i] = c->zn[0][i] + c->acor[i];
  if (!c->acnrmcur) c->acnrm = wrms(N, c->acor, c->ewt);
  return 0;
}

static int handle_nflag(cvo* c, int* nflag, double saved_t, int* ncf) {
  if (*nflag == 0) return DO_ERROR_TEST;
  c->ncfn++;
  cv_restore(c, saved_t);
  if (*nflag < 0) return *nflag;
  (*ncf)++;
  c->etamax = 1.0;
  if (fabs(c->h) <= c->hmin * ONEPSM || *ncf == c->maxncf) return CVO_CONV_FAILURE;
  c->eta = fmax(ETACF, c->hmin / fabs(c->h));
  *nflag = PREV_CONV_FAIL;
  cv_rescale(c);
  return PREDICT_AGAIN;
}

static int do_error_test(cvo* c, int* nflag, double saved_t, int* nef, double* dsmp) {
  double dsm = c->acnrm * c->tq[2];
  *dsmp = dsm;
  if (dsm <= 1.0) return 0;
  (*nef)++; c->netf++; *nflag = PREV_ERR_FAIL;
  cv_restore(c, saved_t);
  if (fabs(c->h) <= c->hmin * ONEPSM || *nef == c->maxnef) return CVO_ERR_FAILURE;
  c->etamax = 1.0;
  if (*nef <= MXNEF1) {
    c->eta = 1.0 / (pow(BIAS2 * dsm, 1.0 / c->L) + ADDON);
    c->eta = fmax(ETAMIN, fmax(c->eta, c->hmin / fabs(c->h)));
    if (*nef >= SMALL_NEF) c->eta = fmin(c->eta, ETAMXF);
    cv_rescale(c);
    return TRY_AGAIN;
  }
  if (c->q > 1) {
    c->eta = fmax(ETAMIN, c->hmin / fabs(c->h));
    cv_adjust_order(c, -1);
    c->L = c->q; c->q--; c->qwait = c->L;
    cv_rescale(c);
    return TRY_AGAIN;
  }
  c->eta = fmax(ETAMIN, c->hmin / fabs(c->h));
  c->h *= c->eta; c->next_h = c->h; c->hscale = c->h; c->qwait = LONG_WAIT;
  int r = c->f(c->tn, c->zn[0], c->tempv, c->user); c->nfe++;
  if (r) return -8;
  for (int i = 0; i < c->N; ++i) c->zn[1][i] = c->h * c->tempv[i];
  return TRY_AGAIN;
}

static void complete_step(cvo* c) {
  int q = c->q, N = c->N;
  c->nst++; c->hu = c->h; c->qu = q;
  for (int i = q; i >= 2; --i) c->tau[i] = c->tau[i - 1];
  if (q == 1 && c->nst > 1) c->tau[2] = c->tau[1];
  c->tau[1] = c->h;
  for (int j = 0; j <= q; ++j) for (int i = 0; i < N; ++i) c->zn[j][i] += c->l[j] * c->acor[i];
  c->qwait--;
  if (c->qwait == 1 && q != c->qmax) { memcpy(c->zn[c->qmax], c->acor, N * 8); c->saved_tq5 = c->tq[5]; }
}

static void set_eta(cvo* c) {
  if (c->eta < THRESH) { c->eta = 1.0; c->hprime = c->h; }
  else {
    c->eta = fmin(c->eta, c->etamax);
    c->eta /= fmax(1.0, fabs(c->h) * c->hmax_inv * c->eta);
    c->hprime = c->h * c->eta;
  }
}
static void prepare_next_step(cvo* c, double dsm) {
  int N = c->N;
  if (c->etamax == 1.0) { c->qwait = c->qwait > 2 ? c->qwait : 2; c->qprime = c->q; c->hprime = c->h; c->eta = 1.0; return; }
  c->etaq = 1.0 / (pow(BIAS2 * dsm, 1.0 / c->L) + ADDON);
  if (c->qwait != 0) { c->eta = c->etaq; c->qprime = c->q; set_eta(c); return; }
  c->qwait = 2;
  /* cvComputeEtaqm1 */
  c->etaqm1 = 0.0;
  if (c->q > 1) {
    double ddn = wrms(N, c->zn[c->q], c->ewt) * c->tq[1];
    c->etaqm1 = 1.0 / (pow(BIAS1 * ddn, 1.0 / c->q) + ADDON);
  }
  /* cvComputeEtaqp1 */
  c->etaqp1 = 0.0;
  if (c->q != c->qmax && c->saved_tq5 != 0.0) {
    double cquot = (c->tq[5] / c->saved_tq5) * pow(c->h / c->tau[2], (double)c->L);
    for (int i = 0; i < N; ++i) c->tempv[i] = -cquot * c->zn[c->qmax][i] + c->acor[i];
    double dup = wrms(N, c->tempv, c->ewt) * c->tq[3];
    c->etaqp1 = 1.0 / (pow(BIAS3 * dup, 1.0 / (c->L + 1)) + ADDON);
  }
  /* cvChooseEta */
  double etam = fmax(c->etaqm1, fmax(c->etaq, c->etaqp1));
  if (etam < THRESH) { c->eta = 1.0; c->qprime = c->q; }
  else if (etam == c->etaq) { c->eta = c->etaq; c->qprime = c->q; }
  else if (etam == c->etaqm1) { c->eta = c->etaqm1; c->qprime = c->q - 1; }
  else { c->eta = c->etaqp1; c->qprime = c->q + 1; memcpy(c->zn[c->qmax], c->acor, N * 8); }
  set_eta(c);
}

static void watch_step(cvo* c) {
  if (c->watch_comp < 0 || c->watch_done) return;
  int k = c->watch_comp;
  if (!(c->zn[0][k] >= c->watch_thr)) return;
  /* root of the Nordsieck interpolant T(s)=sum zn[j][k] s^j on s in [-1,0] by bisection */
  double lo = -1.0, hi = 0.0;
  for (int it = 0; it < 60; ++it) {
    double s = 0.5 * (lo + hi), v = c->zn[c->q][k];
    for (int j = c->q - 1; j >= 0; --j) v = v * s + c->zn[j][k];
    if (v >= c->watch_thr) hi = s; else lo = s;
  }
  c->watch_time = c->tn + hi * c->h;
  c->watch_done = 1;
}

static int cv_step(cvo* c) {
  double saved_t = c->tn, dsm = 0.0;
  int ncf = 0, nef = 0, nflag = FIRST_CALL, kflag, eflag;
  if (c->nst > 0 && c->hprime != c->h) cv_adjust_params(c);
  for (;;) {
    cv_predict(c);
    cv_set_bdf(c);
    nflag = cv_nls(c, nflag);
    kflag = handle_nflag(c, &nflag, saved_t, &ncf);
    if (kflag == PREDICT_AGAIN) continue;
    if (kflag != DO_ERROR_TEST) return kflag;
    eflag = do_error_test(c, &nflag, saved_t, &nef, &dsm);
    if (eflag == TRY_AGAIN) continue;
    if (eflag != 0) return eflag;
    break;
  }
  complete_step(c);
  prepare_next_step(c, dsm);
  c->etamax = (c->nst <= SMALL_NST) ? ETAMX2 : ETAMX3;
  for (int i = 0; i < c->N; ++i) c->acor[i] *= c->tq[2];
  watch_step(c);
  return 0;
}

static void get_dky0(cvo* c, double t, double* out) {
  double s = (t - c->tn) / c->h;
  for (int j = c->q; j >= 0; --j) {
    if (j == c->q) for (int i = 0; i < c->N; ++i) out[i] = c->zn[j][i];
    else for (int i = 0; i < c->N; ++i) out[i] = c->zn[j][i] + s * out[i];
  }
}

int cvo_integrate(cvo* c, double tout, double* yout, double* tret) {
  int N = c->N, istate;
  if (c->nst == 0) {
    *tret = c->tn;
    if (ewt_set(c, c->zn[0])) return CVO_ILL_INPUT;
    if (c->f(c->tn, c->zn[0], c->zn[1], c->user)) return -8;
    c->nfe++;
    if (c->tstopset && (c->tstop - c->tn) * (tout - c->tn) <= 0.0) return CVO_ILL_INPUT;
    c->h = c->hin;
    if (c->h != 0.0 && (tout - c->tn) * c->h < 0.0) return CVO_ILL_INPUT;
    if (c->h == 0.0) {
      double tout_hin = tout;
      if (c->tstopset && (tout - c->tn) * (tout - c->tstop) > 0.0) tout_hin = c->tstop;
      int r = cv_hin(c, tout_hin);
      if (r) return r;
    }
    double rh = fabs(c->h) * c->hmax_inv;
    if (rh > 1.0) c->h /= rh;
    if (fabs(c->h) < c->hmin) c->h *= c->hmin / fabs(c->h);
    if (c->tstopset && (c->tn + c->h - c->tstop) * c->h > 0.0) c->h = (c->tstop - c->tn) * (1.0 - 4.0 * UROUND);
    c->hscale = c->h; c->h0u = c->h; c->hprime = c->h;
    for (int i = 0; i < N; ++i) c->zn[1][i] *= c->h;
    if (c->watch_comp >= 0 && !c->watch_done && c->zn[0][c->watch_comp] >= c->watch_thr) { c->watch_time = c->tn; c->watch_done = 1; }
  }
  if (c->nst > 0) {
    double troundoff = FUZZ_FACTOR * UROUND * (fabs(c->tn) + fabs(c->h));
    if ((c->tn - tout) * c->h >= 0.0) { *tret = tout; get_dky0(c, tout, yout); return CVO_SUCCESS; }
    if (c->tstopset) {
      if (fabs(c->tn - c->tstop) <= troundoff) { get_dky0(c, c->tstop, yout); *tret = c->tstop; c->tstopset = 0; return CVO_TSTOP_RETURN; }
      if ((c->tn + c->hprime - c->tstop) * c->h > 0.0) { c->hprime = (c->tstop - c->tn) * (1.0 - 4.0 * UROUND); c->eta = c->hprime / c->h; }
    }
  }
  long nstloc = 0;
  for (;;) {
    c->next_h = c->h; c->next_q = c->q;
    if (c->nst > 0 && ewt_set(c, c->zn[0])) { istate = CVO_ILL_INPUT; *tret = c->tn; memcpy(yout, c->zn[0], N * 8); break; }
    if (c->mxstep > 0 && nstloc >= c->mxstep) { istate = CVO_TOO_MUCH_WORK; *tret = c->tn; memcpy(yout, c->zn[0], N * 8); break; }
    double nrm = wrms(N, c->zn[0], c->ewt);
    if (UROUND * nrm > 1.0) { istate = CVO_TOO_MUCH_ACC; *tret = c->tn; memcpy(yout, c->zn[0], N * 8); break; }
    int kflag = cv_step(c);
    if (kflag != 0) { istate = kflag; *tret = c->tn; memcpy(yout, c->zn[0], N * 8); break; }
    nstloc++;
    if ((c->tn - tout) * c->h >= 0.0) { istate = CVO_SUCCESS; *tret = tout; get_dky0(c, tout, yout); c->next_h = c->hprime; c->next_q = c->qprime; break; }
    if (c->tstopset) {
      double troundoff = FUZZ_FACTOR * UROUND * (fabs(c->tn) + fabs(c->h));
      if (fabs(c->tn - c->tstop) <= troundoff) { get_dky0(c, c->tstop, yout); *tret = c->tstop; c->tstopset = 0; istate = CVO_TSTOP_RETURN; break; }
      if ((c->tn + c->hprime - c->tstop) * c->h > 0.0) { c->hprime = (c->tstop - c->tn) * (1.0 - 4.0 * UROUND); c->eta = c->hprime / c->h; }
    }
  }
  return istate;
}
