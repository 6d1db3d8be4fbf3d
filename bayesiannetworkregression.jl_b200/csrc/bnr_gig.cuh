// Generalized-inverse-Gaussian sampler GIG(lambda, chi, psi), density ~ x^(lambda-1) exp(-(chi/x + psi x)/2).
// Behavioural restatement of the reference's sample_gig (src/gig.jl:8-42, a Hormann-Leydold GIGrvg
// translation): same branch thresholds, same set-up constants, same order of uniform consumption, so that
// with an injected uniform stream the accepted draw matches the reference to rounding.
#pragma once
#include "bnr_rng.cuh"

namespace bnr {

constexpr int GIG_MAX_ATTEMPTS = 100000;

__device__ inline double gig_mode(double lam, double om) {  // src/gig.jl:170-176
  if (lam >= 1.0) return (sqrt((lam - 1.0) * (lam - 1.0) + om * om) + lam - 1.0) / om;
  return om / (sqrt((1.0 - lam) * (1.0 - lam) + om * om) + (1.0 - lam));
}

// ratio-of-uniforms with mode shift (src/gig.jl:44-78)
__device__ inline double gig_rou_shift(double lam, double om, double alpha, DrawStream& st, bool& capped) {
  const double t = 0.5 * (lam - 1.0), s = 0.25 * om;
  const double xm = gig_mode(lam, om);
  const double nc = t * log(xm) - s * (xm + 1.0 / xm);
  const double a = -(2.0 * (lam + 1.0) / om + xm);
  const double b = 2.0 * (lam - 1.0) * xm / om - 1.0;
  const double c = xm;
  const double p = b - a * a / 3.0;
  const double q = 2.0 * a * a * a / 27.0 - a * b / 3.0 + c;
  const double fi = acos(-q / (2.0 * sqrt(-p * p * p / 27.0)));
  const double fak = 2.0 * sqrt(-p / 3.0);
  const double y1 = fak * cos(fi / 3.0) - a / 3.0;
  const double y2 = fak * cos(fi / 3.0 + 4.0 / 3.0 * 3.141592653589793) - a / 3.0;
  const double uplus = (y1 - xm) * exp(t * log(y1) - s * (y1 + 1.0 / y1) - nc);
  const double uminus = (y2 - xm) * exp(t * log(y2) - s * (y2 + 1.0 / y2) - nc);
  for (int it = 0; it < GIG_MAX_ATTEMPTS; ++it) {
    const double U = uminus + st.uniform() * (uplus - uminus);
    const double V = st.uniform();
    const double X = U / V + xm;
    if (X > 0.0 && log(V) <= t * log(X) - s * (X + 1.0 / X) - nc) return alpha * X;
    if (st.exhausted) break;
  }
  capped = true;
  return alpha * xm;
}

// ratio-of-uniforms without shift (src/gig.jl:80-100)
__device__ inline double gig_rou_noshift(double lam, double om, double alpha, DrawStream& st, bool& capped) {
  const double t = 0.5 * (lam - 1.0), s = 0.25 * om;
  const double xm = gig_mode(lam, om);
  const double nc = t * log(xm) - s * (xm + 1.0 / xm);
  const double ym = ((lam + 1.0) + sqrt((lam + 1.0) * (lam + 1.0) + om * om)) / om;
  const double um = exp(0.5 * (lam + 1.0) * log(ym) - s * (ym + 1.0 / ym) - nc);
  for (int it = 0; it < GIG_MAX_ATTEMPTS; ++it) {
    const double U = um * st.uniform();
    const double V = st.uniform();
    const double X = U / V;
    if (log(V) <= t * log(X) - s * (X + 1.0 / X) - nc) return alpha * X;
    if (st.exhausted) break;
  }
  capped = true;
  return alpha * xm;
}

// three-part envelope for the log-concave region (src/gig.jl:102-168), lambda in (0,1)
__device__ inline double gig_concave(double lam, double om, double alpha, DrawStream& st, bool& capped) {
  const double xm = gig_mode(lam, om);
  const double x0 = om / (1.0 - lam);
  const double k0 = exp((lam - 1.0) * log(xm) - 0.5 * om * (xm + 1.0 / xm));
  const double A1 = k0 * x0;
  double k1, k2, A2, A3;
  if (x0 >= 2.0 / om) {
    k1 = 0.0; A2 = 0.0;
    k2 = pow(x0, lam - 1.0);
    A3 = k2 * 2.0 * exp(-om * x0 / 2.0) / om;
  } else {
    k1 = exp(-om);
    A2 = (lam == 0.0) ? k1 * log(2.0 / (om * om)) : k1 / lam * (pow(2.0 / om, lam) - pow(x0, lam));
    k2 = pow(2.0 / om, lam - 1.0);
    A3 = k2 * 2.0 * exp(-1.0) / om;
  }
  const double Atot = A1 + A2 + A3;
  for (int it = 0; it < GIG_MAX_ATTEMPTS; ++it) {
    double V = Atot * st.uniform();
    double X, hx;
    if (V <= A1) {
      X = x0 * V / A1; hx = k0;
    } else {
      V -= A1;
      if (V <= A2) {
        if (lam == 0.0) { X = om * exp(exp(om) * V); hx = k1 / X; }
        else { X = pow(pow(x0, lam) + lam / k1 * V, 1.0 / lam); hx = k1 * pow(X, lam - 1.0); }
      } else {
        V -= A2;
        const double a = (x0 > 2.0 / om) ? x0 : 2.0 / om;
        X = -2.0 / om * log(exp(-om / 2.0 * a) - om / (2.0 * k2) * V);
        hx = k2 * exp(-om / 2.0 * X);
      }
    }
    const double U = st.uniform() * hx;
    if (log(U) <= (lam - 1.0) * log(X) - om / 2.0 * (X + 1.0 / X)) return alpha * X;
    if (st.exhausted) break;
  }
  capped = true;
  return alpha * xm;
}

// dispatcher for lambda >= 0 (the sampler is only ever called with lambda = 1/2, src/gibbs.jl:116-118)
__device__ inline double sample_gig(double lam, double chi, double psi, DrawStream& st, bool& capped) {
  const double eps10 = 10.0 * 2.220446049250313e-16;
  if (chi < eps10) return st.gamma(lam) * (psi / 2.0);          // reference's own scale convention
  if (psi < eps10) return 1.0 / (st.gamma(lam) * (chi / 2.0));  // (src/gig.jl:15-26)
  const double alpha = sqrt(chi / psi);
  const double om = sqrt(psi * chi);
  if (lam > 2.0 || om > 3.0) return gig_rou_shift(lam, om, alpha, st, capped);
  if (lam >= 1.0 - 2.25 * om * om || om > 0.2) return gig_rou_noshift(lam, om, alpha, st, capped);
  return gig_concave(lam, om, alpha, st, capped);
}

}  // namespace bnr
