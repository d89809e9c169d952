"""Independent high-precision (mpmath, 40 digits) evaluation of the reactor right-hand side, written straight from the
mechanism dump (oracle/mech/*.json, file units) and the formulas SURVEY.md section 8(c) lists for Cantera 2.x
(Phase::setMassFractions, IdealGasPhase, NasaPoly2, GasKinetics::updateROP, Arrhenius, Troe, ThirdBodyCalc) and for the
reactor equations (/root/reference/src/applications/reactors/reactors.cpp:35-57, :88-109, base.cpp:89-97).

Test infrastructure only.  It shares no code with oracle/ct_kinetics.c (own stoichiometry tables, own unit conversion,
own thermo) — only the named constant set — so agreement to ~1e-14 pins the oracle's double-precision arithmetic far
below the 1e-5 the reference-held golden vectors can resolve.  Inputs are taken as the exact binary doubles.
"""
import json
import os

from mpmath import mp, mpf

mp.dps = 40
_HERE = os.path.dirname(os.path.abspath(__file__))


class MpMechanism:
    def __init__(self, name, constants):
        """constants: dict with gas_constant, one_atm, cal, weights (the named set the oracle uses)."""
        with open(os.path.join(_HERE, "..", "oracle", "mech", name + ".json")) as f:
            src = json.load(f)
        u = src["units"]
        assert (u["length"], u["quantity"], u["act_energy"], u["time"]) == ("cm", "mol", "cal/mol", "s")
        self.R = mpf(constants["gas_constant"])
        self.p_ref = mpf(constants["one_atm"])
        self.names = [s["name"] for s in src["species"]]
        self.ix = {n: k for k, n in enumerate(self.names)}
        self.K = len(self.names)
        self.W = [sum(mpf(n) * mpf(constants["weights"][el]) for el, n in s["atoms"]) for s in src["species"]]
        self.lo = [[mpf(a) for a in s["nasa_lo"]["a"]] for s in src["species"]]
        self.hi = [[mpf(a) for a in s["nasa_hi"]["a"]] for s in src["species"]]
        self.Tmid = [mpf(s["nasa_lo"]["T"][1]) for s in src["species"]]
        cal_to_K = mpf(constants["cal"]) * 1000 / self.R  # cal/mol -> J/kmol -> K
        self.rx = []
        for r in src["reactions"]:
            nu_r = {}
            nu_p = {}
            for n, c in r["reactants"]:
                nu_r[self.ix[n]] = nu_r.get(self.ix[n], 0) + int(c)
            for n, c in r["products"]:
                nu_p[self.ix[n]] = nu_p.get(self.ix[n], 0) + int(c)
            m = sum(nu_r.values()) + (1 if r["kind"] == "three_body" else 0)
            # 1 mol/cm^3 = 1e3 kmol/m^3: A_SI = A * (1e-3)^(m-1)
            rec = dict(kind=r["kind"], rev=bool(r["reversible"]), nu_r=nu_r, nu_p=nu_p,
                       A=mpf(r["kf"][0]) * mpf(10) ** (-3 * (m - 1)), b=mpf(r["kf"][1]), E=mpf(r["kf"][2]) * cal_to_K,
                       eff={self.ix[n]: mpf(e) for n, e in r["efficiencies"]})
            if r["kind"] == "falloff":
                rec.update(A0=mpf(r["kf0"][0]) * mpf(10) ** (-3 * m), b0=mpf(r["kf0"][1]), E0=mpf(r["kf0"][2]) * cal_to_K, troe=None)
                if r["falloff"] is not None:
                    f = r["falloff"]
                    rec["troe"] = (mpf(f["A"]), mpf(f["T3"]), mpf(f["T1"]), None if f["T2"] is None else mpf(f["T2"]))
            self.rx.append(rec)

    def thermo(self, T):
        cp, h, s = [], [], []
        for k in range(self.K):
            a = self.lo[k] if T <= self.Tmid[k] else self.hi[k]
            cp.append(a[0] + a[1] * T + a[2] * T ** 2 + a[3] * T ** 3 + a[4] * T ** 4)
            h.append(a[0] + a[1] * T / 2 + a[2] * T ** 2 / 3 + a[3] * T ** 3 / 4 + a[4] * T ** 4 / 5 + a[5] / T)
            s.append(a[0] * mp.log(T) + a[1] * T + a[2] * T ** 2 / 2 + a[3] * T ** 3 / 3 + a[4] * T ** 4 / 4 + a[6])
        return cp, h, s

    def rhs(self, rtype, T0, p0, Y0, state):
        """rtype 0..3 (Const_Pres, Const_Vol, Const_Temp_Pres, Const_Temp_Vol).  Returns ydot, wdot, qf, qr as mpf lists."""
        T0, p0 = mpf(float(T0)), mpf(float(p0))
        energy = rtype <= 1
        T = mpf(float(state[0])) if energy else T0
        Yin = [mpf(float(v)) for v in (state[1:] if energy else state)]

        def normalise(Y):
            Yc = [y if y > 0 else mpf(0) for y in Y]
            tot = sum(Yc)
            Yn = [y / tot for y in Yc]
            return Yn, 1 / sum(y / w for y, w in zip(Yn, self.W))
        Y, mmw = normalise(Yin)
        if rtype in (0, 2):
            p = p0
            rho = p * mmw / (self.R * T)
        else:  # density frozen at construction from (T0, p0, Y0)
            Yn0, mmw0 = normalise([mpf(float(v)) for v in Y0])
            rho = p0 * mmw0 / (self.R * T0)
            p = rho * self.R * T / mmw
        c = [rho * y / w for y, w in zip(Y, self.W)]
        ctot = sum(c)
        cp, h, s = self.thermo(T)
        g = [hk - sk for hk, sk in zip(h, s)]  # g0/RT at p_ref
        lnT = mp.log(T)
        ln_c0 = mp.log(self.p_ref / (self.R * T))
        qf, qr = [], []
        for r in self.rx:
            kf = r["A"] * mp.exp(r["b"] * lnT - r["E"] / T)
            M = None
            if r["kind"] != "elementary":
                M = ctot + sum((e - 1) * c[k] for k, e in r["eff"].items())
            if r["kind"] == "three_body":
                kf = kf * M
            elif r["kind"] == "falloff":
                k0 = r["A0"] * mp.exp(r["b0"] * lnT - r["E0"] / T)
                pr = M * k0 / (kf + mpf("1e-300"))
                F = mpf(1)
                if r["troe"] is not None:
                    a, T3, T1, T2 = r["troe"]
                    Fc = (1 - a) * mp.exp(-T / T3) + a * mp.exp(-T / T1)
                    if T2 is not None and T2 != 0:
                        Fc += mp.exp(-T2 / T)
                    lgFc = mp.log10(max(Fc, mpf("1e-300")))
                    lpr = mp.log10(max(pr, mpf("1e-300")))
                    cc = -mpf("0.4") - mpf("0.67") * lgFc
                    nn = mpf("0.75") - mpf("1.27") * lgFc
                    f1 = (lpr + cc) / (nn - mpf("0.14") * (lpr + cc))
                    F = mpf(10) ** (lgFc / (1 + f1 * f1))
                kf = kf * pr / (1 + pr) * F
            f = kf
            for k, nu in r["nu_r"].items():
                f *= c[k] ** nu
            b = mpf(0)
            if r["rev"]:
                dn = sum(r["nu_p"].values()) - sum(r["nu_r"].values())
                dg = sum(nu * g[k] for k, nu in r["nu_p"].items()) - sum(nu * g[k] for k, nu in r["nu_r"].items())
                b = kf * mp.exp(dg - dn * ln_c0)  # k_f / Kc
                for k, nu in r["nu_p"].items():
                    b *= c[k] ** nu
            qf.append(f)
            qr.append(b)
        wdot = [mpf(0)] * self.K
        for r, f, b in zip(self.rx, qf, qr):
            for k, nu in r["nu_p"].items():
                wdot[k] += nu * (f - b)
            for k, nu in r["nu_r"].items():
                wdot[k] -= nu * (f - b)
        ydot = [w * wk / rho for w, wk in zip(wdot, self.W)]
        self.last = dict(rho=rho, T=T)
        if energy:
            X = [y * mmw / w for y, w in zip(Y, self.W)]
            cp_mass = self.R * sum(x * cpk for x, cpk in zip(X, cp)) / mmw
            if rtype == 0:
                Tdot = -sum(self.R * T * hk * w for hk, w in zip(h, wdot)) / (rho * cp_mass)
                self.last.update(e_molar=[self.R * T * hk for hk in h], rho_c=rho * cp_mass)
            else:
                cv_mass = cp_mass - self.R / mmw
                Tdot = -sum(self.R * T * (hk - 1) * w for hk, w in zip(h, wdot)) / (rho * cv_mass)
                self.last.update(e_molar=[self.R * T * (hk - 1) for hk in h], rho_c=rho * cv_mass)
            ydot = [Tdot] + ydot
        return ydot, wdot, qf, qr
