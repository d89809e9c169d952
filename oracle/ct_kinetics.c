/* ORACLE — test infrastructure, NOT product code (see ct_oracle.h).
 *
 * Restates, in plain scalar C, the Cantera 2.x arithmetic that the reference
 * reaches through
 *   m_thermo.setState_TPY / setState_TRY / density / pressure
 *   m_kinetics.getNetProductionRates
 *   m_thermo.getPartialMolarEnthalpies / getPartialMolarIntEnergies / cp_mass / cv_mass
 * (/root/reference/src/applications/reactors/reactors.cpp:35-57, :88-109, :144-152,
 *  :186-194; /root/reference/src/applications/reactors/base.cpp:62-69, :89-97, :192-200).
 * Cantera itself is an un-vendored dependency; the algorithm restated is that
 * of Cantera 2.x Phase::setMassFractions, IdealGasPhase, NasaPoly2,
 * GasKinetics::updateROP, Arrhenius::updateRC, Troe::F, ThirdBodyCalc::update.
 */
#include "ct_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SMALL_NUMBER 1.e-300
#define BIG_NUMBER 1.e300

struct omech {
  int K, R;
  double Rgas, p_ref;
  double *W, *nlo, *nhi, *Tmid;
  int *rtype, *rev;
  double *A, *b, *EaR, *A0, *b0, *EaR0;
  int* ftype;
  double* troe;
  int *rp, *rs, *pp, *ps, *ep, *es;
  double* ev;
  double* dn; /* sum nu'' - sum nu' (third body excluded) */
};

static double* dcopy(const double* s, size_t n) {
  double* d = (double*)malloc((n ? n : 1) * sizeof(double));
  if (n) memcpy(d, s, n * sizeof(double));
  return d;
}
static int* icopy(const int* s, size_t n) {
  int* d = (int*)malloc((n ? n : 1) * sizeof(int));
  if (n) memcpy(d, s, n * sizeof(int));
  return d;
}

omech* om_create(int K, int R, double gas_constant, double p_ref, const double* W,
                 const double* nasa_lo, const double* nasa_hi, const double* Tmid, const int* rtype,
                 const int* reversible, const double* A, const double* b, const double* EaR,
                 const double* A0, const double* b0, const double* EaR0, const int* ftype,
                 const double* troe, const int* reac_ptr, const int* reac_sp, const int* prod_ptr,
                 const int* prod_sp, const int* eff_ptr, const int* eff_sp, const double* eff_val) {
  omech* m = (omech*)calloc(1, sizeof(omech));
  m->K = K; m->R = R; m->Rgas = gas_constant; m->p_ref = p_ref;
  m->W = dcopy(W, K); m->nlo = dcopy(nasa_lo, 7 * K); m->nhi = dcopy(nasa_hi, 7 * K);
  m->Tmid = dcopy(Tmid, K);
  m->rtype = icopy(rtype, R); m->rev = icopy(reversible, R);
  m->A = dcopy(A, R); m->b = dcopy(b, R); m->EaR = dcopy(EaR, R);
  m->A0 = dcopy(A0, R); m->b0 = dcopy(b0, R); m->EaR0 = dcopy(EaR0, R);
  m->ftype = icopy(ftype, R); m->troe = dcopy(troe, 4 * R);
  m->rp = icopy(reac_ptr, R + 1); m->rs = icopy(reac_sp, reac_ptr[R]);
  m->pp = icopy(prod_ptr, R + 1); m->ps = icopy(prod_sp, prod_ptr[R]);
  m->ep = icopy(eff_ptr, R + 1); m->es = icopy(eff_sp, eff_ptr[R]);
  m->ev = dcopy(eff_val, eff_ptr[R]);
  m->dn = (double*)malloc(R * sizeof(double));
  for (int i = 0; i < R; ++i)
    m->dn[i] = (double)(prod_ptr[i + 1] - prod_ptr[i]) - (double)(reac_ptr[i + 1] - reac_ptr[i]);
  return m;
}

void om_destroy(omech* m) {
  if (!m) return;
  free(m->W); free(m->nlo); free(m->nhi); free(m->Tmid); free(m->rtype); free(m->rev);
  free(m->A); free(m->b); free(m->EaR); free(m->A0); free(m->b0); free(m->EaR0);
  free(m->ftype); free(m->troe); free(m->rp); free(m->rs); free(m->pp); free(m->ps);
  free(m->ep); free(m->es); free(m->ev); free(m->dn); free(m);
}

int om_nspecies(const omech* m) { return m->K; }
int om_nreactions(const omech* m) { return m->R; }

/* NasaPoly1::updateProperties with NasaPoly2's `T <= midT -> low range`. */
void om_species_thermo(const omech* m, double T, double* cp_R, double* h_RT, double* s_R) {
  double T2 = T * T, T3 = T2 * T, T4 = T3 * T, rT = 1.0 / T, lnT = log(T);
  for (int k = 0; k < m->K; ++k) {
    const double* a = (T <= m->Tmid[k]) ? m->nlo + 7 * k : m->nhi + 7 * k;
    double ct0 = a[0], ct1 = a[1] * T, ct2 = a[2] * T2, ct3 = a[3] * T3, ct4 = a[4] * T4;
    cp_R[k] = ct0 + ct1 + ct2 + ct3 + ct4;
    h_RT[k] = ct0 + 0.5 * ct1 + 1.0 / 3.0 * ct2 + 0.25 * ct3 + 0.2 * ct4 + a[5] * rT;
    s_R[k] = ct0 * lnT + ct1 + 0.5 * ct2 + 1.0 / 3.0 * ct3 + 0.25 * ct4 + a[6];
  }
}

/* ---------------------------------------------------------------------- */
struct oreactor {
  const omech* m;
  int type;
  double T0, p0;
  double* Y0;
  double rho, p, T; /* current thermo state */
  double rho_init;  /* density of the initial state (OR_TABLE_V: rho(t) = rho_init V(t0) / V(t)) */
  double mmw;
  /* work */
  double *ym, *conc, *cp_R, *h_RT, *s_R, *grt, *wdot, *qf, *qr, *rkc, *energy;
  int nt;
  double *tab_t, *tab_T, *tab_p;
};

/* Phase::setMassFractions: clip negatives, normalise, Y/W, mean molecular weight. */
static void set_mass_fractions(oreactor* r, const double* y) {
  const omech* m = r->m;
  double norm = 0.0;
  for (int k = 0; k < m->K; ++k) {
    double v = y[k] > 0.0 ? y[k] : 0.0;
    r->ym[k] = v;
    norm += v;
  }
  double sum = 0.0;
  for (int k = 0; k < m->K; ++k) {
    double yk = r->ym[k] / norm; /* scale(m_y, 1/norm) */
    r->ym[k] = yk / m->W[k];     /* m_ym = m_y * m_rmolwts */
    sum += r->ym[k];
  }
  r->mmw = 1.0 / sum;
}

/* setState_TPY: rho = p*mmw/(R*T) (IdealGasPhase::setPressure) */
static void set_state_TPY(oreactor* r, double T, double p, const double* y) {
  set_mass_fractions(r, y);
  r->T = T;
  r->rho = p * r->mmw / (r->m->Rgas * T);
  r->p = p;
}
/* setState_TRY; pressure() = R * molarDensity * T */
static void set_state_TRY(oreactor* r, double T, double rho, const double* y) {
  set_mass_fractions(r, y);
  r->T = T;
  r->rho = rho;
  r->p = r->m->Rgas * (rho / r->mmw) * T;
}

/* GasKinetics::updateROP (+update_rates_T, update_rates_C, updateKc, processFalloffReactions). */
static void update_rop(oreactor* r) {
  const omech* m = r->m;
  const int K = m->K, R = m->R;
  const double T = r->T, logT = log(T), recipT = 1.0 / T;
  const double RT = m->Rgas * T;
  om_species_thermo(m, T, r->cp_R, r->h_RT, r->s_R);
  /* concentrations: c_k = rho * Y_k/W_k; ctot = molarDensity */
  double ctot = r->rho / r->mmw;
  for (int k = 0; k < K; ++k) r->conc[k] = r->rho * r->ym[k];
  /* getStandardChemPotentials: mu0_k = RT*(g0_RT + ln(P/Pref)) */
  double tmp = log(r->p / m->p_ref);
  for (int k = 0; k < K; ++k) r->grt[k] = RT * ((r->h_RT[k] - r->s_R[k]) + tmp);
  double logStandConc = log(r->p / RT); /* IdealGasPhase::standardConcentration = P/RT */
  double rrt = 1.0 / RT;
  for (int i = 0; i < R; ++i) {
    /* rate coefficient (Arrhenius::updateRC: A*exp(b*logT - E*recipT)) */
    double kf = m->A[i] * exp(m->b[i] * logT - m->EaR[i] * recipT);
    double concm = 0.0;
    if (m->rtype[i] != 0) {
      /* ThirdBodyCalc::update: default*ctot + sum (eff-default)*c */
      concm = ctot;
      for (int e = m->ep[i]; e < m->ep[i + 1]; ++e) concm += (m->ev[e] - 1.0) * r->conc[m->es[e]];
    }
    double ropf;
    if (m->rtype[i] == 1) {
      ropf = kf * concm;
    } else if (m->rtype[i] == 2) {
      double k0 = m->A0[i] * exp(m->b0[i] * logT - m->EaR0[i] * recipT);
      double pr = concm * k0 / (kf + SMALL_NUMBER);
      double F = 1.0;
      if (m->ftype[i] == 1) {
        const double* tp = m->troe + 4 * i; /* A, T3, T1, T2 */
        double Fcent = (1.0 - tp[0]) * exp(-T * (1.0 / tp[1])) + tp[0] * exp(-T * (1.0 / tp[2]));
        if (tp[3] != 0.0) Fcent += exp(-tp[3] / T);
        double lgFc = log10(Fcent > SMALL_NUMBER ? Fcent : SMALL_NUMBER);
        double lpr = log10(pr > SMALL_NUMBER ? pr : SMALL_NUMBER);
        double cc = -0.4 - 0.67 * lgFc;
        double nn = 0.75 - 1.27 * lgFc;
        double f1 = (lpr + cc) / (nn - 0.14 * (lpr + cc));
        double lgf = lgFc / (1.0 + f1 * f1);
        F = pow(10.0, lgf);
      }
      pr *= F / (1.0 + pr); /* FalloffMgr::pr_to_falloff */
      ropf = pr * kf;       /* pr[i] *= m_rfn_high[i] */
    } else {
      ropf = kf;
    }
    /* reciprocal equilibrium constant (GasKinetics::updateKc) */
    double rkc = 0.0;
    if (m->rev[i]) {
      double dg = 0.0;
      for (int s = m->pp[i]; s < m->pp[i + 1]; ++s) dg += r->grt[m->ps[s]];
      for (int s = m->rp[i]; s < m->rp[i + 1]; ++s) dg -= r->grt[m->rs[s]];
      rkc = exp(dg * rrt - m->dn[i] * logStandConc);
      if (rkc > BIG_NUMBER) rkc = BIG_NUMBER;
    }
    double ropr = ropf * rkc;
    for (int s = m->rp[i]; s < m->rp[i + 1]; ++s) ropf *= r->conc[m->rs[s]];
    for (int s = m->pp[i]; s < m->pp[i + 1]; ++s) ropr *= r->conc[m->ps[s]];
    r->qf[i] = ropf;
    r->qr[i] = ropr;
  }
  /* getNetProductionRates: products increment, reactants decrement net ROP */
  for (int k = 0; k < K; ++k) r->wdot[k] = 0.0;
  for (int i = 0; i < R; ++i) {
    double net = r->qf[i] - r->qr[i];
    for (int s = m->pp[i]; s < m->pp[i + 1]; ++s) r->wdot[m->ps[s]] += net;
    for (int s = m->rp[i]; s < m->rp[i + 1]; ++s) r->wdot[m->rs[s]] -= net;
  }
}

static double mean_X(const oreactor* r, const double* q) {
  double s = 0.0;
  for (int k = 0; k < r->m->K; ++k) s += r->ym[k] * q[k];
  return r->mmw * s;
}

static void table_lookup(const oreactor* r, double t, double* T, double* p) {
  int n = r->nt;
  if (n <= 0) { *T = r->T0; *p = r->p0; return; }
  if (t <= r->tab_t[0]) { *T = r->tab_T[0]; *p = r->tab_p[0]; return; }
  if (t >= r->tab_t[n - 1]) { *T = r->tab_T[n - 1]; *p = r->tab_p[n - 1]; return; }
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    int mid = (lo + hi) / 2;
    if (r->tab_t[mid] <= t) lo = mid; else hi = mid;
  }
  double w = (t - r->tab_t[lo]) / (r->tab_t[hi] - r->tab_t[lo]);
  *T = r->tab_T[lo] + w * (r->tab_T[hi] - r->tab_T[lo]);
  *p = r->tab_p[lo] + w * (r->tab_p[hi] - r->tab_p[lo]);
}

oreactor* or_create(const omech* m, int type, double T, double p, const double* Y) {
  oreactor* r = (oreactor*)calloc(1, sizeof(oreactor));
  int K = m->K, R = m->R;
  r->m = m; r->type = type; r->T0 = T; r->p0 = p;
  r->Y0 = dcopy(Y, K);
  r->ym = (double*)calloc(K, 8); r->conc = (double*)calloc(K, 8); r->cp_R = (double*)calloc(K, 8);
  r->h_RT = (double*)calloc(K, 8); r->s_R = (double*)calloc(K, 8); r->grt = (double*)calloc(K, 8);
  r->wdot = (double*)calloc(K, 8); r->energy = (double*)calloc(K, 8);
  r->qf = (double*)calloc(R, 8); r->qr = (double*)calloc(R, 8); r->rkc = (double*)calloc(R, 8);
  /* constructors: ConstVol/ConstTempVol fix rho from (T0,p0,Y0) (reactors.cpp:83-85,:175-180);
     ConstPres/ConstTempPres keep p (reactors.cpp:35-42,:133-141). */
  set_state_TPY(r, T, p, Y);
  r->rho_init = r->rho;
  return r;
}
void or_set_table(oreactor* r, int nt, const double* t, const double* T, const double* p) {
  free(r->tab_t); free(r->tab_T); free(r->tab_p);
  r->nt = nt; r->tab_t = dcopy(t, nt); r->tab_T = dcopy(T, nt); r->tab_p = dcopy(p, nt);
}
void or_destroy(oreactor* r) {
  if (!r) return;
  free(r->Y0); free(r->ym); free(r->conc); free(r->cp_R); free(r->h_RT); free(r->s_R);
  free(r->grt); free(r->wdot); free(r->energy); free(r->qf); free(r->qr); free(r->rkc);
  free(r->tab_t); free(r->tab_T); free(r->tab_p); free(r);
}
static int has_energy(int type) { return type <= OR_CONST_VOL || type == OR_TABLE_V; }
int or_nstate(const oreactor* r) { return has_energy(r->type) ? r->m->K + 1 : r->m->K; }
void or_initial_state(const oreactor* r, double* y) {
  int K = r->m->K;
  if (has_energy(r->type)) { y[0] = r->T0; memcpy(y + 1, r->Y0, K * 8); }
  else memcpy(y, r->Y0, K * 8);
}
double or_density(const oreactor* r) { return r->rho; }
double or_pressure(const oreactor* r) { return r->p; }

int or_rhs(oreactor* r, double t, const double* y, double* ydot) {
  const omech* m = r->m;
  const int K = m->K;
  double* dY;
  double vdv = 0.0; /* OR_TABLE_V: (dV/dt) / V */
  switch (r->type) {
    case OR_CONST_PRES: set_state_TPY(r, y[0], r->p0, y + 1); dY = ydot + 1; break;
    case OR_CONST_VOL: {
      /* rho frozen at construction */
      double rho0 = r->rho;
      set_state_TRY(r, y[0], rho0, y + 1); dY = ydot + 1; break;
    }
    case OR_CONST_TEMP_PRES: set_state_TPY(r, r->T0, r->p0, y); dY = ydot; break;
    case OR_CONST_TEMP_VOL: { double rho0 = r->rho; set_state_TRY(r, r->T0, rho0, y); dY = ydot; break; }
    case OR_TABLE_TP: { double T, p; table_lookup(r, t, &T, &p); set_state_TPY(r, T, p, y); dY = ydot; break; }
    case OR_TABLE_V: {
      /* closed system of fixed mass in the prescribed volume: rho = rho_init V(t0) / V(t) */
      double V, Vd;
      table_lookup(r, t, &V, &Vd);
      const double V0 = r->nt > 0 ? r->tab_T[0] : 1.0;
      if (r->nt <= 0) { V = 1.0; Vd = 0.0; }
      set_state_TRY(r, y[0], r->rho_init * (V0 / V), y + 1); dY = ydot + 1;
      vdv = Vd / V;
      break;
    }
    default: return -1;
  }
  update_rop(r);
  /* Base::SpeciesDerivative (base.cpp:89-97) */
  for (int k = 0; k < K; ++k) dY[k] = r->wdot[k] * m->W[k] / r->rho;
  if (r->type == OR_CONST_PRES) {
    /* ConstPres::EnergyDerivative (reactors.cpp:44-57) */
    double RT = m->Rgas * r->T, ip = 0.0;
    for (int k = 0; k < K; ++k) { r->energy[k] = r->h_RT[k] * RT; }
    for (int k = 0; k < K; ++k) ip += r->energy[k] * r->wdot[k];
    double cp_mass = m->Rgas * mean_X(r, r->cp_R) / r->mmw;
    ydot[0] = (-ip / r->rho / cp_mass);
  } else if (r->type == OR_CONST_VOL || r->type == OR_TABLE_V) {
    /* ConstVol::EnergyDerivative (reactors.cpp:96-109) */
    double RT = m->Rgas * r->T, ip = 0.0;
    for (int k = 0; k < K; ++k) { r->energy[k] = RT * (r->h_RT[k] - 1.0); }
    for (int k = 0; k < K; ++k) ip += r->energy[k] * r->wdot[k];
    double cp_mole = m->Rgas * mean_X(r, r->cp_R);
    double cv_mass = (cp_mole - m->Rgas) / r->mmw;
    ydot[0] = (-ip / r->rho / cv_mass);
    if (r->type == OR_TABLE_V) ydot[0] -= (r->p * vdv) / r->rho / cv_mass; /* boundary work -p dV */
  }
  return 0;
}

void or_last_rates(const oreactor* r, double* wdot, double* qf, double* qr) {
  if (wdot) memcpy(wdot, r->wdot, r->m->K * 8);
  if (qf) memcpy(qf, r->qf, r->m->R * 8);
  if (qr) memcpy(qr, r->qr, r->m->R * 8);
}

void om_mixture_props(const omech* m, double T, double p, const double* Y, double* rho,
                      double* h_mole, double* s_mole, double* cp_mole, double* mmw) {
  oreactor* r = or_create(m, OR_CONST_PRES, T, p, Y);
  om_species_thermo(m, T, r->cp_R, r->h_RT, r->s_R);
  if (rho) *rho = r->rho;
  if (mmw) *mmw = r->mmw;
  if (h_mole) *h_mole = m->Rgas * T * mean_X(r, r->h_RT);
  if (cp_mole) *cp_mole = m->Rgas * mean_X(r, r->cp_R);
  if (s_mole) {
    /* IdealGasPhase::entropy_mole = R*(mean_X(s0_R) - sum_xlogx - log(p/pref)) */
    double sxlx = 0.0;
    for (int k = 0; k < m->K; ++k) {
      double x = r->ym[k] * r->mmw;
      if (x > SMALL_NUMBER) sxlx += x * log(x);
    }
    *s_mole = m->Rgas * (mean_X(r, r->s_R) - sxlx - log(p / m->p_ref));
  }
  or_destroy(r);
}

int or_rhs_batch(const omech* m, int type, long n, const double* T0, const double* p0,
                 const double* Y0, const double* states, double* ydot, double* wdot, double* qf,
                 double* qr) {
  int K = m->K, R = m->R;
  for (long i = 0; i < n; ++i) {
    oreactor* r = or_create(m, type, T0[i], p0[i], Y0 + i * K);
    int N = or_nstate(r);
    or_rhs(r, 0.0, states + i * N, ydot + i * N);
    or_last_rates(r, wdot ? wdot + i * K : NULL, qf ? qf + i * R : NULL, qr ? qr + i * R : NULL);
    or_destroy(r);
  }
  return 0;
}
