// Mechanism compiler back end: MechanismDef (file units) -> CompiledMech
// (Cantera's kmol-m-s-K-J system, device reaction order, one flat blob that
// the kernels stage into shared memory).  See DESIGN.md "Mechanism blob".
#include <algorithm>
#include <cmath>
#include <cstring>

#include "mechanism.hpp"

namespace ctb {

const char* const kDefaultConstants = "cantera-2.6";

const ConstantSet& constant_set(const std::string& name) {
  // Cantera's constants changed between releases and the reference does not
  // pin one.  "cantera-2.6": CODATA-2018 R and abridged IUPAC
  // weights (2.5/2.6); "cantera-2.4": the older table.
  static const ConstantSet sets[] = {
      {"cantera-2.6", 8314.46261815324, 101325.0, 4.184, {{"H", 1.008}, {"C", 12.011}, {"N", 14.007}, {"O", 15.999}, {"Ar", 39.948}}},
      {"cantera-2.4", 8314.4621, 101325.0, 4.184, {{"H", 1.00794}, {"C", 12.0107}, {"N", 14.0067}, {"O", 15.9994}, {"Ar", 39.948}}},
  };
  for (auto& s : sets) if (s.name == name) return s;
  throw MechError("unknown constant set '" + name + "' (known: cantera-2.6, cantera-2.4)");
}

namespace {
template <class T>
int append(std::vector<unsigned char>& blob, const std::vector<T>& v) {
  while (blob.size() % 16) blob.push_back(0);
  int off = (int)blob.size();
  const unsigned char* p = reinterpret_cast<const unsigned char*>(v.data());
  blob.insert(blob.end(), p, p + v.size() * sizeof(T));
  return off;
}
std::string canon_element(const std::string& e) {
  std::string o = e;
  for (size_t i = 0; i < o.size(); ++i) o[i] = i == 0 ? (char)std::toupper((unsigned char)o[i]) : (char)std::tolower((unsigned char)o[i]);
  return o;
}
}  // namespace

CompiledMech compile_mechanism(const MechanismDef& D, const std::string& constants_name) {
  const std::string cname = constants_name.empty() ? std::string(kDefaultConstants) : constants_name;
  const ConstantSet& cs = constant_set(cname);
  if (D.u_length != "cm" || D.u_quantity != "mol" || D.u_act_energy != "cal/mol" || D.u_time != "s")
    throw MechError("unsupported CTI unit system (only length=cm, time=s, quantity=mol, act_energy=cal/mol)");
  CompiledMech C;
  C.constants = cs.name;
  C.gas_constant = cs.gas_constant;
  C.p_ref = cs.one_atm;
  const int K = C.K = (int)D.species.size();
  const int R = C.R = (int)D.reactions.size();
  if (K < 1 || K > 250) throw MechError("species count out of range (1..250)");
  if (R < 1 || R > 65000) throw MechError("reaction count out of range");
  C.KP = ((K + 1 + 31) / 32) * 32;  // room for the "no species" sentinel slot K
  C.elements = D.elements;
  std::map<std::string, int> sp_index;
  for (int k = 0; k < K; ++k) {
    if (sp_index.count(D.species[k].name)) throw MechError("duplicate species " + D.species[k].name);
    sp_index[D.species[k].name] = k;
    C.species_names.push_back(D.species[k].name);
  }
  // molecular weights
  C.W.assign(K, 0.0);
  C.natoms.assign((size_t)K * D.elements.size(), 0.0);
  for (int k = 0; k < K; ++k) {
    for (auto& a : D.species[k].atoms) {
      std::string el = canon_element(a.first);
      auto it = cs.weights.find(el);
      if (it == cs.weights.end()) throw MechError("no atomic weight for element '" + a.first + "'");
      C.W[k] += a.second * it->second;
      for (size_t e = 0; e < D.elements.size(); ++e)
        if (canon_element(D.elements[e]) == el) C.natoms[k * D.elements.size() + e] = a.second;
    }
  }
  // device order: falloff | three-body | elementary with a temperature-dependent k_f | elementary with b = 0 and
  // Ea = 0 (k_f = A: Cantera's A exp(0) is exactly A, so the lane-passes over that tail skip the exponential);
  // stable within each class.  The temperature-dependent region is padded with constant reactions (evaluated as
  // A exp(0), exact) to a multiple of 32, so the constant tail starts on a lane-pass boundary: n_exp_end.
  {
    std::vector<int> cls_list[4];
    for (int i = 0; i < R; ++i) {
      const ReactionDef& r = D.reactions[i];
      const bool constant = r.kind == RxnKind::Elementary && r.kf[1] == 0.0 && r.kf[2] == 0.0;
      cls_list[r.kind == RxnKind::Falloff ? 0 : r.kind == RxnKind::ThreeBody ? 1 : constant ? 3 : 2].push_back(i);
    }
    size_t moved = 0;
    while (cls_list[2].size() % 32 != 0 && moved < cls_list[3].size()) cls_list[2].push_back(cls_list[3][moved++]);
    cls_list[3].erase(cls_list[3].begin(), cls_list[3].begin() + moved);
    for (int c = 0; c < 4; ++c) for (int i : cls_list[c]) C.perm.push_back(i);
    C.n_const = (int)cls_list[3].size();
    C.n_exp_end = (int)(cls_list[0].size() + cls_list[1].size() + cls_list[2].size());
  }
  C.inv_perm.assign(R, 0);
  for (int d = 0; d < R; ++d) C.inv_perm[C.perm[d]] = d;
  // Kinetics::reactionString: Cantera rebuilds the text from its species->coefficient std::map, i.e. species in
  // ASCII order on each side (the reference asserts that form, tests/src/chemistry.cpp:287, :299).
  auto side_string = [](const std::vector<std::pair<std::string, int>>& terms, RxnKind kind) {
    std::map<std::string, int> acc;
    for (auto& t : terms) acc[t.first] += t.second;
    std::string s;
    for (auto& kv : acc) {
      if (!s.empty()) s += " + ";
      if (kv.second != 1) s += std::to_string(kv.second) + " ";
      s += kv.first;
    }
    if (kind == RxnKind::ThreeBody) s += " + M";
    if (kind == RxnKind::Falloff) s += " (+M)";
    return s;
  };
  for (auto& r : D.reactions) {
    C.equations.push_back(side_string(r.reactants, r.kind) + (r.reversible ? " <=> " : " => ") + side_string(r.products, r.kind));
    if (r.kind == RxnKind::Falloff) C.n_fall++;
    if (r.kind == RxnKind::ThreeBody) C.n_3b++;
  }
  const int nf = C.n_fall, n3 = C.n_3b;
  const double to_EaR = cs.cal_to_J * 1000.0 / cs.gas_constant;  // cal/mol -> J/kmol -> K

  std::vector<double> rW(K), Tmid(K), nasa((size_t)2 * 7 * C.KP, 0.0);
  for (int k = 0; k < K; ++k) {
    rW[k] = 1.0 / C.W[k];
    Tmid[k] = D.species[k].Tmid;
    for (int c = 0; c < 7; ++c) {
      nasa[(size_t)(0 * 7 + c) * C.KP + k] = D.species[k].a_lo[c];
      nasa[(size_t)(1 * 7 + c) * C.KP + k] = D.species[k].a_hi[c];
    }
  }
  std::vector<double> A(R), b(R), E(R), A0(nf), b0(nf), E0(nf), ta(nf), rt3(nf), rt1(nf), t2(nf), eff_val;
  std::vector<int> eff_ptr(nf + n3 + 1, 0);
  std::vector<unsigned short> eff_o;  // collider species as byte offsets into the per-warp {c, 1/E} pair array (16 * k)
  std::vector<unsigned char> ftype(nf, 0), rev(R), reac((size_t)3 * R, (unsigned char)K), prod((size_t)3 * R, (unsigned char)K);
  std::vector<signed char> dn(R);
  // species -> reaction lists with net stoichiometric coefficient
  std::vector<std::vector<std::pair<int, int>>> sp_list(K);
  for (int d = 0; d < R; ++d) {
    const ReactionDef& r = D.reactions[C.perm[d]];
    int order = 0, nprod = 0;
    std::map<int, int> net;
    int slot = 0;
    for (auto& t : r.reactants) {
      auto it = sp_index.find(t.first);
      if (it == sp_index.end()) throw MechError("undeclared species '" + t.first + "' in: " + r.equation);
      for (int c = 0; c < t.second; ++c) {
        if (slot >= 3) throw MechError("more than 3 reactant molecules in: " + r.equation);
        reac[(size_t)slot++ * R + d] = (unsigned char)it->second;
      }
      order += t.second;
      net[it->second] -= t.second;
    }
    slot = 0;
    for (auto& t : r.products) {
      auto it = sp_index.find(t.first);
      if (it == sp_index.end()) throw MechError("undeclared species '" + t.first + "' in: " + r.equation);
      for (int c = 0; c < t.second; ++c) {
        if (slot >= 3) throw MechError("more than 3 product molecules in: " + r.equation);
        prod[(size_t)slot++ * R + d] = (unsigned char)it->second;
      }
      nprod += t.second;
      net[it->second] += t.second;
    }
    for (auto& kv : net) if (kv.second != 0) sp_list[kv.first].emplace_back(d, kv.second);
    dn[d] = (signed char)(nprod - order);
    rev[d] = r.reversible ? 1 : 0;
    // 1 mol/cm^3 = 1e3 kmol/m^3  =>  A_SI = A * (1e-3)^(m-1), m = molecularity incl. M
    int m = order + (r.kind == RxnKind::ThreeBody ? 1 : 0);
    A[d] = r.kf[0] * std::pow(1e-3, m - 1);
    b[d] = r.kf[1];
    E[d] = r.kf[2] * to_EaR;
    if (r.kind == RxnKind::Falloff) {
      A0[d] = r.kf0[0] * std::pow(1e-3, order);
      b0[d] = r.kf0[1];
      E0[d] = r.kf0[2] * to_EaR;
      ftype[d] = r.troe ? 1 : 0;
      ta[d] = r.troe_p[0];
      rt3[d] = r.troe ? 1.0 / r.troe_p[1] : 0.0;
      rt1[d] = r.troe ? 1.0 / r.troe_p[2] : 0.0;
      t2[d] = r.troe_p[3];
    }
    if (r.kind != RxnKind::Elementary) {
      for (auto& e : r.efficiencies) {
        auto it = sp_index.find(e.first);
        if (it == sp_index.end()) throw MechError("undeclared collider '" + e.first + "' in: " + r.equation);
        eff_o.push_back((unsigned short)(16 * it->second));
        eff_val.push_back(e.second - 1.0);
      }
      eff_ptr[d + 1] = (int)eff_o.size();
    } else if (!r.efficiencies.empty()) {
      throw MechError("efficiencies on an elementary reaction: " + r.equation);
    }
  }
  std::vector<int> sp_ptr(K + 1, 0);
  std::vector<unsigned short> sp_rxn;
  std::vector<signed char> sp_nu;
  for (int k = 0; k < K; ++k) {
    for (auto& e : sp_list[k]) { sp_rxn.push_back((unsigned short)e.first); sp_nu.push_back((signed char)e.second); }
    sp_ptr[k + 1] = (int)sp_rxn.size();
  }
  auto& o = C.off;
  // Range check for the per-species form of 1/Kc (kernels.cuh: recip_kc): the product of three exp(+-mu0/RT) must stay
  // inside the fp64 range, i.e. |g0/RT| + |ln(p/p_ref)| (<= 10 up to 2e4 atm) below 230 for every species.  g0/RT falls
  // in magnitude with T, so the lowest temperature that passes is searched from 400 K upwards.
  {
    auto g_rt = [](const double* a, double T) {
      return a[0] * (1.0 - std::log(T)) - a[1] * T / 2.0 - a[2] * T * T / 6.0 - a[3] * T * T * T / 12.0 - a[4] * T * T * T * T / 20.0 + a[5] / T - a[6];
    };
    double Tmin = 400.0;
    for (; Tmin < 3000.0; Tmin += 100.0) {
      double worst = 0.0;
      for (int k = 0; k < K; ++k) {
        const SpeciesDef& sp = D.species[k];
        worst = std::max(worst, std::fabs(g_rt(Tmin <= sp.Tmid ? sp.a_lo : sp.a_hi, Tmin)));
      }
      if (worst + 10.0 < 230.0) break;
    }
    C.kc_fast_Tmin = Tmin;
  }
  o.W = append(C.blob, C.W); o.rW = append(C.blob, rW); o.Tmid = append(C.blob, Tmid); o.nasa = append(C.blob, nasa);
  o.A = append(C.blob, A); o.b = append(C.blob, b); o.E = append(C.blob, E);
  o.A0 = append(C.blob, A0); o.b0 = append(C.blob, b0); o.E0 = append(C.blob, E0);
  o.troe_a = append(C.blob, ta); o.troe_rt3 = append(C.blob, rt3); o.troe_rt1 = append(C.blob, rt1); o.troe_t2 = append(C.blob, t2);
  o.eff_val = append(C.blob, eff_val); o.eff_ptr = append(C.blob, eff_ptr); o.sp_ptr = append(C.blob, sp_ptr);
  o.sp_rxn = append(C.blob, sp_rxn); o.sp_nu = append(C.blob, sp_nu);
  o.eff_o = append(C.blob, eff_o); o.ftype = append(C.blob, ftype);
  // packed reaction records, 16 bytes each (one LDS.128): six 16-bit byte offsets, pre-scaled for the arrays the
  // right-hand side indexes with them -- reactant slots 16 * k into the per-warp {c_k, 1/E_k} pair array, product slots
  // 8 * k into the c_k E_k array (K = "no species": a slot that holds 1.0) -- then 8 * index into the per-warp table of
  // C^-dn factors {C^2, C, 1, 1/C, 1/C^2, 0} (index dn + 2; 5 = irreversible: no reverse rate), then a spare.
  std::vector<unsigned short> rxo((size_t)8 * R, 0);
  for (int d = 0; d < R; ++d) {
    for (int sl = 0; sl < 3; ++sl) rxo[(size_t)8 * d + sl] = (unsigned short)(16 * reac[(size_t)sl * R + d]);
    for (int sl = 0; sl < 3; ++sl) rxo[(size_t)8 * d + 3 + sl] = (unsigned short)(8 * prod[(size_t)sl * R + d]);
    if (dn[d] < -2 || dn[d] > 2) throw MechError("change of mole number outside -2..2 in: " + D.reactions[C.perm[d]].equation);
    rxo[(size_t)8 * d + 6] = (unsigned short)(8 * (rev[d] ? dn[d] + 2 : 5));
  }
  o.rxo = append(C.blob, rxo);
  // balanced gather tables for wdot: the species-sorted entry list (|nu| = 2 entries duplicated) is cut at species
  // boundaries and at multiples of ceil(nnz/32); lane l owns the pieces that start inside its chunk.
  {
    std::vector<unsigned short> ent, pstart, pnplus;
    std::vector<unsigned char> lpp(33, 0), spp(K + 1, 0);
    std::vector<std::vector<std::pair<int, int>>> flat(K);  // per species: (reaction, sign) with duplicates
    size_t nnz = 0;
    for (int k = 0; k < K; ++k)
      for (auto& e : sp_list[k])
        for (int c = 0; c < std::abs(e.second); ++c) { flat[k].emplace_back(e.first, e.second > 0 ? 1 : -1); ++nnz; }
    const size_t chunk = std::max<size_t>(1, (nnz + 31) / 32);
    size_t pos = 0;
    std::vector<int> piece_lane;
    for (int k = 0; k < K; ++k) {
      spp[k] = (unsigned char)pstart.size();
      size_t i = 0;
      while (i < flat[k].size()) {
        const size_t room = chunk - (pos % chunk);
        const size_t len = std::min(room, flat[k].size() - i);
        // entries are byte offsets into the per-warp q[] array (reaction index * 8), stored in PAIRS (one 32-bit load
        // per two terms); an odd run is padded with the offset of q[R], a slot that is never written and stays zero
        pstart.push_back((unsigned short)(ent.size() / 2));
        piece_lane.push_back((int)(pos / chunk));
        const unsigned short zero_slot = (unsigned short)(8 * R);
        unsigned short np = 0;
        for (size_t j = i; j < i + len; ++j) if (flat[k][j].second > 0) { ent.push_back((unsigned short)(8 * flat[k][j].first)); ++np; }
        if (np & 1) { ent.push_back(zero_slot); ++np; }
        unsigned short nm = 0;
        for (size_t j = i; j < i + len; ++j) if (flat[k][j].second < 0) { ent.push_back((unsigned short)(8 * flat[k][j].first)); ++nm; }
        if (nm & 1) ent.push_back(zero_slot);
        pnplus.push_back((unsigned short)(np / 2));
        i += len; pos += len;
      }
    }
    spp[K] = (unsigned char)pstart.size();
    const int P = (int)pstart.size();
    if (P > 128 || P > 250) throw MechError("too many gather pieces for the shared partial buffer");
    pstart.push_back((unsigned short)(ent.size() / 2));
    if (8 * (R + 1) > 65535) throw MechError("too many reactions for 16-bit gather offsets");
    for (int l = 0; l <= 32; ++l) {
      int first = P;
      for (int pc = P - 1; pc >= 0; --pc) if (piece_lane[pc] >= l) first = pc;
      lpp[l] = (unsigned char)first;
    }
    o.ent = append(C.blob, ent); o.pstart = append(C.blob, pstart); o.pnplus = append(C.blob, pnplus);
    o.lpp = append(C.blob, lpp); o.spp = append(C.blob, spp); o.n_pieces = P;
  }
  while (C.blob.size() % 16) C.blob.push_back(0);
  o.n_eff = (int)eff_o.size(); o.n_nnz = (int)sp_rxn.size(); o.total = (int)C.blob.size();

  // ---- algorithmic flop model (FMA = 2, each fp64 exp/log/log10/pow = 20) ----
  int n_rev = 0, n_troe = 0, n_stoich = 0;
  for (auto& r : D.reactions) {
    n_rev += r.reversible;
    n_troe += (r.kind == RxnKind::Falloff && r.troe);
    for (auto& t : r.reactants) n_stoich += t.second;
    for (auto& t : r.products) n_stoich += t.second;
  }
  const int n_arr = R + nf;
  C.n_transcendental = 1 /*lnT*/ + 2 /*ln p/pref, ln c0*/ + n_arr + n_rev + n_troe * 6;
  double arith = 8.0 * K            /* clip, normalise, Y/W, c */
                 + 31.0 * K         /* NASA-7 cp,h,s,g */
                 + 6.0 * K          /* energy sums, dY/dt */
                 + 5.0 * n_arr      /* Arrhenius argument + scale */
                 + 10.0 * n_rev     /* delta g gather + Kc argument */
                 + 2.0 * (double)eff_o.size() + 2.0 * (nf + n3) + 24.0 * nf /* third body, falloff algebra */
                 + 6.0 * R          /* concentration products, net */
                 + 2.0 * n_stoich;  /* wdot scatter */
  C.flops_rhs = arith + 20.0 * C.n_transcendental;
  const int N = K + 1;
  // analytic d(wdot)/dc assembly: each (species,reaction) entry touches <=6 slots (mul+fma = 3 flops),
  // third-body rows K per entry, plus chain rule to (T,Y) K*K*3 and energy row; T column = one extra RHS.
  C.flops_jac = C.flops_rhs + 6.0 * 3.0 * sp_rxn.size() + 3.0 * (double)K * K + 6.0 * K + 2.0 * K * (nf + n3) + 2.0 * (double)K * K;
  C.flops_lu = 2.0 / 3.0 * (double)N * N * N;
  C.flops_solve = 2.0 * (double)N * N;
  return C;
}

}  // namespace ctb
