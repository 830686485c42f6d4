// celeste_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (double precision, C++17, no dependencies) of the reference's
// per-source ELBO hot path, written AS THE REFERENCE WRITES IT: dense 44x44
// SensitiveFloat algebra, per-pixel scan over all S sources, per-component
// chain rule, same loop order and the same quirks.  It is the parity oracle for
// the CUDA library and the "port" CPU baseline of bench.py.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load it; the product (celeste.jl_b200/) never does.
//
// PARITY UNPINNED: the reference (Julia 0.6 + un-vendored packages) cannot run
// in this environment and its test-suite holds no stored numbers for this path
// (SURVEY.md 8c); the cubic B-spline evaluation rule of the un-vendored
// Interpolations.jl (REQUIRE:21, no version pin) is restated from its published
// algorithm.  The oracle is pinned instead by (a) the reference's one
// closed-form check (get_bvn_cov, test/test_elbo.jl:45-61), (b) the reference's
// own self-consistency strategy -- hand derivatives == automatic
// differentiation of the same value code (test/test_elbo.jl:223-301), here
// against an independent torch.float64 autograd model (tests/ad_model.py) --
// and (c) the structural property tests of test/test_elbo.jl.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/src).  Indices here are 0-based; comments give the 1-based
// reference indices where that helps.

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/celeste_cuda.h"

namespace {

constexpr int P = CELESTE_NUM_PARAMS;  // length(CanonicalParams) param_set.jl:107
constexpr int NB = 5;                  // NUM_BANDS
constexpr int NT = 2;                  // NUM_SOURCE_TYPES light_source_model.jl:7
constexpr int BP = 10;                 // length(BrightnessParams) param_set.jl:74

// ---- param_set.jl:76-103 canonical ids (0-based) ---------------------------
constexpr int ID_POS0 = 0, ID_POS1 = 1;
constexpr int ID_FRAC_DEV = 2, ID_AXIS_RATIO = 3, ID_ANGLE = 4, ID_RADIUS = 5;
inline int id_flux_loc(int i) { return 6 + i; }
inline int id_flux_scale(int i) { return 8 + i; }
inline int id_color_mean(int m, int i) { return 10 + m + 4 * i; }
inline int id_color_var(int m, int i) { return 18 + m + 4 * i; }
inline int id_is_star(int i) { return 26 + i; }
// BrightnessParams param_set.jl:63-70 (0-based): flux_loc 0, flux_scale 1, color_mean 2..5, color_var 6..9
constexpr int BID_FLUX_LOC = 0, BID_FLUX_SCALE = 1;
inline int bid_color_mean(int m) { return 2 + m; }
inline int bid_color_var(int m) { return 6 + m; }

// shape_standard_alignment / brightness_standard_alignment param_set.jl:161-164
struct Align {
    int shape[NT][6];
    int n_shape[NT];
    int bright[NT][BP];
};
const Align& align() {
    static Align a = [] {
        Align r{};
        r.n_shape[0] = 2;
        r.shape[0][0] = ID_POS0;
        r.shape[0][1] = ID_POS1;
        r.n_shape[1] = 6;
        int g[6] = {ID_POS0, ID_POS1, ID_FRAC_DEV, ID_AXIS_RATIO, ID_ANGLE, ID_RADIUS};
        for (int k = 0; k < 6; ++k) r.shape[1][k] = g[k];
        for (int i = 0; i < NT; ++i) {
            r.bright[i][0] = id_flux_loc(i);
            r.bright[i][1] = id_flux_scale(i);
            for (int m = 0; m < 4; ++m) r.bright[i][2 + m] = id_color_mean(m, i);
            for (int m = 0; m < 4; ++m) r.bright[i][6 + m] = id_color_var(m, i);
        }
        return r;
    }();
    return a;
}
// gal_shape_alignment param_set.jl:168: GalaxyShapeParams -> GalaxyPosParams (0-based 3,4,5)
constexpr int GAL_SHAPE_ALIGN[3] = {3, 4, 5};

// ---- light_source_model.jl:45-72 galaxy prototypes --------------------------
struct Prototypes {
    double eta[2][8];
    double nu[2][8];
    int n[2];
};
const Prototypes& prototypes() {
    static Prototypes p = [] {
        Prototypes r{};
        double dev_amp[8] = {4.26347652e-2, 2.40127183e-1, 6.85907632e-1, 1.51937350,
                             2.83627243,    4.46467501,    5.72440830,    5.60989349};
        double dev_var[8] = {2.23759216e-4, 1.00220099e-3, 4.18731126e-3, 1.69432589e-2,
                             6.84850479e-2, 2.87207080e-1, 1.33320254,    8.40215071};
        double exp_amp[6] = {2.34853813e-3, 3.07995260e-2, 2.23364214e-1,
                             1.17949102,    4.33873750,    5.99820770};
        double exp_var[6] = {1.20078965e-3, 8.84526493e-3, 3.91463084e-2,
                             1.39976817e-1, 4.60962500e-1, 1.50159566};
        double sd = 0, se = 0;
        for (double a : dev_amp) sd += a;
        for (double a : exp_amp) se += a;
        const double er[2] = {1.078031, 0.928896};
        r.n[0] = 8;
        r.n[1] = 6;
        for (int j = 0; j < 8; ++j) {
            r.eta[0][j] = dev_amp[j] / sd;
            r.nu[0][j] = dev_var[j] / (er[0] * er[0]);
        }
        for (int j = 0; j < 6; ++j) {
            r.eta[1][j] = exp_amp[j] / se;
            r.nu[1][j] = exp_var[j] / (er[1] * er[1]);
        }
        return r;
    }();
    return p;
}

// ---- SensitiveFloats.jl:23-47 ------------------------------------------------
struct SF {
    double v = 0;
    std::vector<double> d;  // local_P x local_S
    std::vector<double> h;  // (P*S) x (P*S), column-major
    int lP = 0, lS = 0;
    bool has_gradient = false, has_hessian = false;
    SF() = default;
    SF(int lp, int ls, bool g, bool hh) : lP(lp), lS(ls), has_gradient(g), has_hessian(hh) {
        d.assign(g ? (size_t)lp * ls : 0, 0.0);
        size_t hd = hh ? (size_t)lp * ls : 0;
        h.assign(hd * hd, 0.0);
    }
    int hdim() const { return has_hessian ? lP * lS : 0; }
    double& H(int i, int j) { return h[(size_t)i + (size_t)j * hdim()]; }
    double H(int i, int j) const { return h[(size_t)i + (size_t)j * hdim()]; }
};

// SensitiveFloats.jl:83-93
void zero(SF& sf) {
    sf.v = 0;
    if (sf.has_gradient) std::fill(sf.d.begin(), sf.d.end(), 0.0);
    if (sf.has_hessian) std::fill(sf.h.begin(), sf.h.end(), 0.0);
}
// SensitiveFloats.jl:67-75
void set_hess(SF& sf, int i, int j, double v) {
    sf.H(i, j) = v;
    sf.H(j, i) = v;
}

// SensitiveFloats.jl:99-128 (fills the full square, column by column)
void combine_sfs_hessian(const SF& sf1, const SF& sf2, SF& res, const double g_d[2],
                         const double g_h[2][2]) {
    const int p2 = res.hdim();
    for (int ind2 = 0; ind2 < p2; ++ind2) {
        const double sf11_factor = g_h[0][0] * sf1.d[ind2] + g_h[0][1] * sf2.d[ind2];
        const double sf21_factor = g_h[0][1] * sf1.d[ind2] + g_h[1][1] * sf2.d[ind2];
        for (int ind1 = 0; ind1 < p2; ++ind1) {
            // NB: res may alias sf1 (multiply_sfs!); the reference reads h[ind1,ind2] of
            // the operands before overwriting the same entry, and d is updated afterwards.
            const double a = sf1.H(ind1, ind2);
            const double b = sf2.H(ind1, ind2);
            res.H(ind1, ind2) = g_d[0] * a + g_d[1] * b + sf11_factor * sf1.d[ind1] +
                                sf21_factor * sf2.d[ind1];
        }
    }
}
// SensitiveFloats.jl:139-168
void combine_sfs(const SF& sf1, const SF& sf2, SF& res, double v, const double g_d[2],
                 const double g_h[2][2]) {
    if (res.has_hessian) combine_sfs_hessian(sf1, sf2, res, g_d, g_h);
    if (res.has_gradient)
        for (size_t i = 0; i < res.d.size(); ++i) res.d[i] = g_d[0] * sf1.d[i] + g_d[1] * sf2.d[i];
    res.v = v;
}
// SensitiveFloats.jl:171-180
void multiply_sfs(SF& sf1, const SF& sf2) {
    const double v = sf1.v * sf2.v;
    const double g_d[2] = {sf2.v, sf1.v};
    const double g_h[2][2] = {{0, 1}, {1, 0}};
    combine_sfs(sf1, sf2, sf1, v, g_d, g_h);
}
// SensitiveFloats.jl:185-208
void add_scaled_sfs(SF& sf1, const SF& sf2, double scale) {
    sf1.v += scale * sf2.v;
    if (sf1.has_gradient)
        for (size_t i = 0; i < sf1.d.size(); ++i) sf1.d[i] += scale * sf2.d[i];
    if (sf1.has_hessian) {
        const int p2 = sf1.hdim();
        for (int ind2 = 0; ind2 < p2; ++ind2)
            for (int ind1 = 0; ind1 <= ind2; ++ind1) {
                sf1.H(ind1, ind2) += scale * sf2.H(ind1, ind2);
                sf1.H(ind2, ind1) = sf1.H(ind1, ind2);
            }
    }
}
// SensitiveFloats.jl:215-250 (s is 0-based here)
void add_sources_sf(SF& all, const SF& s_sf, int s) {
    all.v += s_sf.v;
    const int Pl = all.lP;
    const int shift = Pl * s;
    if (all.has_gradient)
        for (int i = 0; i < Pl; ++i) all.d[shift + i] = all.d[shift + i] + s_sf.d[i];
    if (all.has_hessian)
        for (int i1 = 0; i1 < Pl; ++i1)
            for (int i2 = 0; i2 < Pl; ++i2) all.H(shift + i2, shift + i1) += s_sf.H(i2, i1);
}

// ---- source_brightness.jl:18-202 ---------------------------------------------
struct SourceBrightness {
    SF E_l_a[NB][NT];
    SF E_ll_a[NB][NT];
};

void source_brightness(const double* vs, bool calc_grad, bool calc_hess, SourceBrightness& sb) {
    for (int i = 0; i < NT; ++i) {
        const double flux_loc = vs[id_flux_loc(i)], flux_scale = vs[id_flux_scale(i)];
        double cm[4], cv[4];
        for (int m = 0; m < 4; ++m) {
            cm[m] = vs[id_color_mean(m, i)];
            cv[m] = vs[id_color_var(m, i)];
        }
        SF(&E)[NB][NT] = sb.E_l_a;
        for (int b = 0; b < NB; ++b) E[b][i] = SF(BP, 1, calc_grad, calc_hess);
        // :45-50 (bands 1-based 3,4,5,2,1 -> 0-based 2,3,4,1,0)
        E[2][i].v = std::exp(flux_loc + 0.5 * flux_scale);
        E[3][i].v = std::exp(cm[2] + .5 * cv[2]);
        E[4][i].v = std::exp(cm[3] + .5 * cv[3]);
        E[1][i].v = std::exp(-cm[1] + .5 * cv[1]);
        E[0][i].v = std::exp(-cm[0] + .5 * cv[0]);
        if (calc_grad) {
            // :52-64
            E[2][i].d[BID_FLUX_LOC] = E[2][i].v;
            E[2][i].d[BID_FLUX_SCALE] = E[2][i].v * .5;
            if (calc_hess) {
                set_hess(E[2][i], BID_FLUX_LOC, BID_FLUX_LOC, E[2][i].v);
                set_hess(E[2][i], BID_FLUX_LOC, BID_FLUX_SCALE, E[2][i].v * 0.5);
                set_hess(E[2][i], BID_FLUX_SCALE, BID_FLUX_SCALE, E[2][i].v * 0.25);
            }
            // lognormal colour factor f = exp(sgn*c + v/2): d/dc = sgn f, d/dv = f/2 (:69-107)
            auto colour = [&](int band, int m, double sgn, int times_band) {
                SF& e = E[band][i];
                e.d[bid_color_mean(m)] = e.v * sgn;
                e.d[bid_color_var(m)] = e.v * .5;
                if (calc_hess) {
                    set_hess(e, bid_color_mean(m), bid_color_mean(m), e.v);
                    set_hess(e, bid_color_mean(m), bid_color_var(m), e.v * (sgn * 0.5));
                    set_hess(e, bid_color_var(m), bid_color_var(m), e.v * 0.25);
                }
                multiply_sfs(e, E[times_band][i]);
            };
            colour(3, 2, 1.0, 2);   // band 4 = band 3 * colour 3   (:69-76)
            colour(4, 3, 1.0, 3);   // band 5 = band 4 * colour 4   (:79-86)
            colour(1, 1, -1.0, 2);  // band 2 = band 3 * colour 2   (:89-96)
            colour(0, 0, -1.0, 1);  // band 1 = band 2 * colour 1   (:99-106)
        } else {
            // :108-113
            E[3][i].v *= E[2][i].v;
            E[4][i].v *= E[3][i].v;
            E[1][i].v *= E[2][i].v;
            E[0][i].v *= E[1][i].v;
        }

        SF(&L)[NB][NT] = sb.E_ll_a;
        for (int b = 0; b < NB; ++b) L[b][i] = SF(BP, 1, calc_grad, calc_hess);
        // :123-127
        L[2][i].v = std::exp(2 * flux_loc + 2 * flux_scale);
        L[3][i].v = std::exp(2 * cm[2] + 2 * cv[2]);
        L[4][i].v = std::exp(2 * cm[3] + 2 * cv[3]);
        L[1][i].v = std::exp(-2 * cm[1] + 2 * cv[1]);
        L[0][i].v = std::exp(-2 * cm[0] + 2 * cv[0]);
        if (calc_grad) {
            // :130-139
            L[2][i].d[BID_FLUX_LOC] = 2 * L[2][i].v;
            L[2][i].d[BID_FLUX_SCALE] = 2 * L[2][i].v;
            if (calc_hess) {
                set_hess(L[2][i], BID_FLUX_LOC, BID_FLUX_LOC, 4.0 * L[2][i].v);
                set_hess(L[2][i], BID_FLUX_LOC, BID_FLUX_SCALE, 4.0 * L[2][i].v);
                set_hess(L[2][i], BID_FLUX_SCALE, BID_FLUX_SCALE, 4.0 * L[2][i].v);
            }
            // f = exp(2 sgn c + 2 v): d/dc = 2 sgn f, d/dv = 2 f, second derivs 4 f (cross 4 sgn f) (:141-193)
            auto colour2 = [&](int band, int m, double sgn, int times_band) {
                SF& e = L[band][i];
                e.d[bid_color_mean(m)] = e.v * (2. * sgn);
                e.d[bid_color_var(m)] = e.v * 2.;
                if (calc_hess) {
                    set_hess(e, bid_color_mean(m), bid_color_mean(m), e.v * 4.0);
                    set_hess(e, bid_color_var(m), bid_color_var(m), e.v * 4.0);
                    set_hess(e, bid_color_mean(m), bid_color_var(m), e.v * (4.0 * sgn));
                }
                multiply_sfs(e, L[times_band][i]);
            };
            colour2(3, 2, 1.0, 2);
            colour2(4, 3, 1.0, 3);
            colour2(1, 1, -1.0, 2);
            colour2(0, 0, -1.0, 1);
        } else {
            L[3][i].v *= L[2][i].v;
            L[4][i].v *= L[3][i].v;
            L[1][i].v *= L[2][i].v;
            L[0][i].v *= L[1][i].v;
        }
    }
}

// ---- BivariateNormals.jl ---------------------------------------------------------
// :29-43
void get_bvn_cov(double ab, double angle, double scale, double out[2][2]) {
    const double cp = std::cos(angle), sp = std::sin(angle);
    const double ab_term = ab * ab - 1;
    const double scale_squared = scale * scale;
    const double off_diag_term = -scale_squared * cp * sp * ab_term;
    out[0][0] = scale_squared * (1 + ab_term * (sp * sp));
    out[0][1] = off_diag_term;
    out[1][0] = off_diag_term;
    out[1][1] = scale_squared * (1 + ab_term * (cp * cp));
}

// :143-191
struct BvnComponent {
    double the_mean[2];
    double precision[2][2];
    double z;
    double dsiginv_dsig[3][3];
    double major_sd;
};
BvnComponent make_bvn(const double mean[2], const double cov[2][2], double weight, bool calc_siginv) {
    BvnComponent b{};
    const double det = cov[0][0] * cov[1][1] - cov[0][1] * cov[1][0];
    const double c = 1 / (std::sqrt(det) * 2 * M_PI);
    b.major_sd = std::sqrt(std::fmax(cov[0][0], cov[1][1]));
    b.the_mean[0] = mean[0];
    b.the_mean[1] = mean[1];
    // StaticArrays inv(::SMatrix{2,2}): adjugate * (1/det)
    const double idet = 1 / det;
    b.precision[0][0] = cov[1][1] * idet;
    b.precision[1][0] = -cov[1][0] * idet;
    b.precision[0][1] = -cov[0][1] * idet;
    b.precision[1][1] = cov[0][0] * idet;
    b.z = c * weight;
    if (calc_siginv) {
        const double(&p)[2][2] = b.precision;
        const double d11 = -p[0][0] * p[0][0];
        const double d12 = -2 * p[0][0] * p[0][1];
        const double d13 = -p[0][1] * p[0][1];
        const double d21 = -p[0][0] * p[1][0];
        const double d22 = -(p[0][0] * p[1][1] + p[0][1] * p[0][1]);
        const double d23 = -p[1][1] * p[0][1];
        const double d32 = -2 * p[1][1] * p[1][0];
        const double d33 = -p[1][1] * p[1][1];
        // :181-183 -- row 3, col 1 re-uses dsiginv_dsig13 (same value as dsiginv_dsig31)
        const double m[3][3] = {{d11, d12, d13}, {d21, d22, d23}, {d13, d32, d33}};
        std::memcpy(b.dsiginv_dsig, m, sizeof m);
    }
    return b;
}

// :50-88
struct BvnDerivs {
    double py1, py2, f_pre;
    double bvn_x_d[2], bvn_sig_d[3];
    double bvn_xx_h[2][2], bvn_xsig_h[2][3], bvn_sigsig_h[3][3];
    double dpy1_dsig[3], dpy2_dsig[3];
    double bvn_u_d[2], bvn_uu_h[2][2], bvn_s_d[3], bvn_ss_h[3][3], bvn_us_h[2][3];
};

// :208-222
void eval_bvn_pdf(BvnDerivs& bd, const BvnComponent& bmc, const double x[2]) {
    bd.py1 = bmc.precision[0][0] * (x[0] - bmc.the_mean[0]) + bmc.precision[0][1] * (x[1] - bmc.the_mean[1]);
    bd.py2 = bmc.precision[1][0] * (x[0] - bmc.the_mean[0]) + bmc.precision[1][1] * (x[1] - bmc.the_mean[1]);
    bd.f_pre = bmc.z * std::exp(-0.5 * ((x[0] - bmc.the_mean[0]) * bd.py1 + (x[1] - bmc.the_mean[1]) * bd.py2));
}

// :240-319
void get_bvn_derivs(BvnDerivs& bd, const BvnComponent& bvn, bool calc_x_hess, bool calc_sigma_hess) {
    bd.bvn_x_d[0] = -bd.py1;
    bd.bvn_x_d[1] = -bd.py2;
    if (calc_x_hess) {
        bd.bvn_xx_h[0][0] = -bvn.precision[0][0];
        bd.bvn_xx_h[1][1] = -bvn.precision[1][1];
        bd.bvn_xx_h[0][1] = bd.bvn_xx_h[1][0] = -bvn.precision[0][1];
    }
    bd.bvn_sig_d[0] = 0.5 * bd.py1 * bd.py1 - 0.5 * bvn.precision[0][0];
    bd.bvn_sig_d[1] = bd.py1 * bd.py2 - bvn.precision[0][1];
    bd.bvn_sig_d[2] = 0.5 * bd.py2 * bd.py2 - 0.5 * bvn.precision[1][1];
    if (calc_sigma_hess) {
        bd.dpy1_dsig[0] = -bd.py1 * bvn.precision[0][0];
        bd.dpy1_dsig[1] = -bd.py2 * bvn.precision[0][0] - bd.py1 * bvn.precision[0][1];
        bd.dpy1_dsig[2] = -bd.py2 * bvn.precision[0][1];
        bd.dpy2_dsig[0] = -bd.py1 * bvn.precision[0][1];
        bd.dpy2_dsig[1] = -bd.py1 * bvn.precision[1][1] - bd.py2 * bvn.precision[0][1];
        bd.dpy2_dsig[2] = -bd.py2 * bvn.precision[1][1];
        for (int s = 0; s < 3; ++s) {
            bd.bvn_sigsig_h[0][s] = bd.py1 * bd.dpy1_dsig[s] - 0.5 * bvn.dsiginv_dsig[0][s];
            bd.bvn_sigsig_h[1][s] = bd.py1 * bd.dpy2_dsig[s] + bd.py2 * bd.dpy1_dsig[s] - bvn.dsiginv_dsig[1][s];
            bd.bvn_sigsig_h[2][s] = bd.py2 * bd.dpy2_dsig[s] - 0.5 * bvn.dsiginv_dsig[2][s];
        }
        for (int x = 0; x < 2; ++x) {
            bd.bvn_xsig_h[x][0] = bd.py1 * bvn.precision[0][x];
            bd.bvn_xsig_h[x][1] = bd.py1 * bvn.precision[1][x] + bd.py2 * bvn.precision[0][x];
            bd.bvn_xsig_h[x][2] = bd.py2 * bvn.precision[1][x];
        }
    }
}

// :331-397
struct GalaxySigmaDerivs {
    double j[3][3];     // [sig][shape]
    double t[3][3][3];  // [sig][shape1][shape2]
};
GalaxySigmaDerivs make_sig_derivs(double gal_angle, double gal_axis_ratio, double gal_radius_px,
                                  const double XiXi[2][2], double nuBar, bool calc_tensor) {
    GalaxySigmaDerivs r{};
    const double cos_sin = std::cos(gal_angle) * std::sin(gal_angle);
    const double sin_sq = std::sin(gal_angle) * std::sin(gal_angle);
    const double cos_sq = std::cos(gal_angle) * std::cos(gal_angle);
    const double rr = gal_radius_px * gal_radius_px;
    double j[3][3];
    const double c1 = 2 * gal_axis_ratio * rr;
    j[0][0] = c1 * sin_sq;
    j[1][0] = c1 * -cos_sin;
    j[2][0] = c1 * cos_sq;
    const double c2 = rr * (gal_axis_ratio * gal_axis_ratio - 1);
    j[0][1] = c2 * (2 * cos_sin);
    j[1][1] = c2 * (sin_sq - cos_sq);
    j[2][1] = c2 * (-2 * cos_sin);
    // XiXi[1], XiXi[2], XiXi[4] (linear, column-major) = (1,1), (2,1), (2,2)
    j[0][2] = 2 * XiXi[0][0] / gal_radius_px;
    j[1][2] = 2 * XiXi[1][0] / gal_radius_px;
    j[2][2] = 2 * XiXi[1][1] / gal_radius_px;
    double t[3][3][3] = {};
    if (calc_tensor) {
        const double ab = gal_axis_ratio;
        // column-major fill order of the 27 literals at :364-390: t[sig, s1, s2]
        t[0][0][0] = sin_sq * 2 * rr;
        t[1][0][0] = -cos_sin * 2 * rr;
        t[2][0][0] = cos_sq * 2 * rr;
        t[0][1][0] = 2 * cos_sin * 2 * rr * ab;
        t[1][1][0] = (sin_sq - cos_sq) * 2 * rr * ab;
        t[2][1][0] = -2 * cos_sin * 2 * rr * ab;
        t[0][2][0] = 2 * j[0][0] / gal_radius_px;
        t[1][2][0] = 2 * j[1][0] / gal_radius_px;
        t[2][2][0] = 2 * j[2][0] / gal_radius_px;
        t[0][0][1] = 2 * cos_sin * 2 * rr * ab;
        t[1][0][1] = (sin_sq - cos_sq) * 2 * rr * ab;
        t[2][0][1] = -2 * cos_sin * 2 * rr * ab;
        t[0][1][1] = (cos_sq - sin_sq) * 2 * rr * (ab * ab - 1);
        t[1][1][1] = 2 * cos_sin * 2 * rr * (ab * ab - 1);
        t[2][1][1] = (sin_sq - cos_sq) * 2 * rr * (ab * ab - 1);
        t[0][2][1] = 2 * j[0][1] / gal_radius_px;
        t[1][2][1] = 2 * j[1][1] / gal_radius_px;
        t[2][2][1] = 2 * j[2][1] / gal_radius_px;
        t[0][0][2] = 2 * j[0][0] / gal_radius_px;
        t[1][0][2] = 2 * j[1][0] / gal_radius_px;
        t[2][0][2] = 2 * j[2][0] / gal_radius_px;
        t[0][1][2] = 2 * j[0][1] / gal_radius_px;
        t[1][1][2] = 2 * j[1][1] / gal_radius_px;
        t[2][1][2] = 2 * j[2][1] / gal_radius_px;
        // XiXi[1 << (k-1)], k = 1..3 -> linear elements 1, 2, 4
        t[0][2][2] = 2 * XiXi[0][0] / rr;
        t[1][2][2] = 2 * XiXi[1][0] / rr;
        t[2][2][2] = 2 * XiXi[1][1] / rr;
    }
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            r.j[a][b] = j[a][b] * nuBar;
            for (int c = 0; c < 3; ++c) r.t[a][b][c] = t[a][b][c] * nuBar;
        }
    return r;
}

// :414-448   wcs_jacobian J[a][b] = J[a,b] (row a, col b)
void transform_bvn_ux_derivs(BvnDerivs& bd, const double J[2][2], bool calc_hess) {
    bd.bvn_u_d[0] = -(bd.bvn_x_d[0] * J[0][0] + bd.bvn_x_d[1] * J[1][0]);
    bd.bvn_u_d[1] = -(bd.bvn_x_d[0] * J[0][1] + bd.bvn_x_d[1] * J[1][1]);
    if (calc_hess) {
        std::memset(bd.bvn_uu_h, 0, sizeof bd.bvn_uu_h);
        for (int x2 = 0; x2 < 2; ++x2)
            for (int x1 = 0; x1 < 2; ++x1)
                for (int u2 = 0; u2 < 2; ++u2) {
                    const double inner = bd.bvn_xx_h[x1][x2] * J[x2][u2];
                    for (int u1 = 0; u1 <= u2; ++u1) bd.bvn_uu_h[u1][u2] += inner * J[x1][u1];
                }
        bd.bvn_uu_h[1][0] = bd.bvn_uu_h[0][1];
    }
}
// :465-529
void transform_bvn_derivs_hessian(BvnDerivs& bd, const GalaxySigmaDerivs& sig_sf, const double J[2][2]) {
    std::memset(bd.bvn_ss_h, 0, sizeof bd.bvn_ss_h);
    std::memset(bd.bvn_us_h, 0, sizeof bd.bvn_us_h);
    for (int s2 = 0; s2 < 3; ++s2)
        for (int s1 = 0; s1 <= s2; ++s1)
            for (int g = 0; g < 3; ++g) bd.bvn_ss_h[s1][s2] += bd.bvn_sig_d[g] * sig_sf.t[g][s1][s2];
    for (int g1 = 0; g1 < 3; ++g1)
        for (int g2 = 0; g2 < 3; ++g2)
            for (int s2 = 0; s2 < 3; ++s2) {
                const double inner = bd.bvn_sigsig_h[g1][g2] * sig_sf.j[g2][s2];
                for (int s1 = 0; s1 <= s2; ++s1) bd.bvn_ss_h[s1][s2] += inner * sig_sf.j[g1][s1];
            }
    for (int s2 = 0; s2 < 3; ++s2)
        for (int s1 = 0; s1 <= s2; ++s1) bd.bvn_ss_h[s2][s1] = bd.bvn_ss_h[s1][s2];
    for (int s = 0; s < 3; ++s)
        for (int u = 0; u < 2; ++u)
            for (int g = 0; g < 3; ++g)
                for (int x = 0; x < 2; ++x)
                    bd.bvn_us_h[u][s] += bd.bvn_xsig_h[x][g] * sig_sf.j[g][s] * (-J[x][u]);
}
// :540-572
void transform_bvn_derivs(BvnDerivs& bd, const GalaxySigmaDerivs& sig_sf, const double J[2][2], bool calc_hess) {
    transform_bvn_ux_derivs(bd, J, calc_hess);
    for (int s = 0; s < 3; ++s) {
        bd.bvn_s_d[s] = 0;
        for (int g = 0; g < 3; ++g) bd.bvn_s_d[s] += bd.bvn_sig_d[g] * sig_sf.j[g][s];
    }
    if (calc_hess) transform_bvn_derivs_hessian(bd, sig_sf, J);
}

// ---- model/fsm_util.jl -------------------------------------------------------------
// :29-65
struct GalaxyCacheComponent {
    double gal_frac_dev_dir;
    double gal_frac_dev_i;
    BvnComponent bmc;
    GalaxySigmaDerivs sig_sf;
};

struct PatchView {  // imaged_sources.jl:60-71 through the flat celeste_patch
    const celeste_patch* p;
    double J[2][2];
};
PatchView view(const celeste_patch* p) {
    PatchView v;
    v.p = p;
    v.J[0][0] = p->wcs_jacobian[0];
    v.J[1][0] = p->wcs_jacobian[1];
    v.J[0][1] = p->wcs_jacobian[2];
    v.J[1][1] = p->wcs_jacobian[3];
    return v;
}
// wcs_utils.jl:14-18
void linear_world_to_pix(const PatchView& pv, const double world[2], double out[2]) {
    const double d0 = world[0] - pv.p->world_center[0], d1 = world[1] - pv.p->world_center[1];
    out[0] = (pv.J[0][0] * d0 + pv.J[0][1] * d1) + pv.p->pixel_center[0];
    out[1] = (pv.J[1][0] * d0 + pv.J[1][1] * d1) + pv.p->pixel_center[1];
}

GalaxyCacheComponent make_gcc(double dir, double frac_i, double etaBar, double nuBar, const double* pc /*7*/,
                              const double pos[2], double ab, double angle, double radius, bool calc_grad,
                              bool calc_hess) {
    GalaxyCacheComponent g{};
    double XiXi[2][2];
    get_bvn_cov(ab, angle, radius, XiXi);
    const double mean_s[2] = {pc[1] + pos[0], pc[2] + pos[1]};
    // tauBar col-major: pc[3]=(1,1) pc[4]=(2,1) pc[5]=(1,2) pc[6]=(2,2)
    const double var_s[2][2] = {{pc[3] + nuBar * XiXi[0][0], pc[5] + nuBar * XiXi[0][1]},
                                {pc[4] + nuBar * XiXi[1][0], pc[6] + nuBar * XiXi[1][1]}};
    const double weight = pc[0] * etaBar;
    g.gal_frac_dev_dir = dir;
    g.gal_frac_dev_i = frac_i;
    g.bmc = make_bvn(mean_s, var_s, weight, calc_grad && calc_hess);
    if (calc_grad) g.sig_sf = make_sig_derivs(angle, ab, radius, XiXi, nuBar, calc_hess);
    return g;
}

// ---- Interpolations.jl BSpline(Cubic(Line())), OnGrid: evaluation rule -----------------
// (un-vendored dependency, restated; see header).  coefs padded by 1 per side.
// value, first and second derivatives of the 1-D cubic B-spline weights at fractional offset fx
inline void cubic_weights(double fx, double w[4], double dw[4], double ddw[4]) {
    const double omf = 1 - fx;
    const double fx_cub = fx * fx * fx, omf_cub = omf * omf * omf;
    w[0] = (1.0 / 6) * omf_cub;
    w[1] = 2.0 / 3 - fx * fx + 0.5 * fx_cub;
    w[2] = 2.0 / 3 - omf * omf + 0.5 * omf_cub;
    w[3] = (1.0 / 6) * fx_cub;
    dw[0] = -0.5 * omf * omf;
    dw[1] = -2 * fx + 1.5 * fx * fx;
    dw[2] = 2 * omf - 1.5 * omf * omf;
    dw[3] = 0.5 * fx * fx;
    ddw[0] = omf;
    ddw[1] = -2 + 3 * fx;
    ddw[2] = -2 + 3 * omf;
    ddw[3] = fx;
}
// itp[x, y] with gradient/Hessian in (x, y); n1,n2 = padded dims
void spline_eval(const double* coefs, int n1, int n2, double x, double y, double& val, double g[2], double H[3]) {
    const int s1 = n1 - 2, s2 = n2 - 2;  // size(itp, d)
    int ix = (int)std::floor(x);
    ix = ix < 1 ? 1 : (ix > s1 - 1 ? s1 - 1 : ix);
    int iy = (int)std::floor(y);
    iy = iy < 1 ? 1 : (iy > s2 - 1 ? s2 - 1 : iy);
    const double fx = x - ix, fy = y - iy;
    double wx[4], dwx[4], ddwx[4], wy[4], dwy[4], ddwy[4];
    cubic_weights(fx, wx, dwx, ddwx);
    cubic_weights(fy, wy, dwy, ddwy);
    // padded 1-based index ix+1 is the "ix" tap; taps ix-1..ix+2 -> 0-based rows ix-1 .. ix+2
    double v = 0, gx = 0, gy = 0, hxx = 0, hxy = 0, hyy = 0;
    for (int b = 0; b < 4; ++b) {
        const double* col = coefs + (size_t)(iy - 1 + b) * n1 + (ix - 1);
        double r = 0, rd = 0, rdd = 0;
        for (int a = 0; a < 4; ++a) {
            r += wx[a] * col[a];
            rd += dwx[a] * col[a];
            rdd += ddwx[a] * col[a];
        }
        v += wy[b] * r;
        gx += wy[b] * rd;
        gy += dwy[b] * r;
        hxx += wy[b] * rdd;
        hxy += dwy[b] * rd;
        hyy += ddwy[b] * r;
    }
    val = v;
    g[0] = gx;
    g[1] = gy;
    H[0] = hxx;
    H[1] = hxy;
    H[2] = hyy;
}

// fsm_util.jl:221-248.  fs0m: SF(2,1).  ForwardDiff.gradient!/hessian! of
// f(pos) = softpluslikeinv(itp[h - m_pos1 + 26, w - m_pos2 + 26]) restated analytically.
void star_light_density(SF& fs0m, const PatchView& pv, int h, int w, const double pos[2], bool is_active) {
    double m_pos[2];
    linear_world_to_pix(pv, pos, m_pos);
    double y, g[2], Hh[3];
    spline_eval(pv.p->itp_coefs, pv.p->itp_dims[0], pv.p->itp_dims[1], h - m_pos[0] + 26, w - m_pos[1] + 26, y, g, Hh);
    // softpluslikeinv(y) = y < 0 ? 1e-3exp(y) : 1e-3(y + 1)   (:222)
    double s0, s1, s2;
    if (y < 0) {
        s0 = 1e-3 * std::exp(y);
        s1 = s0;
        s2 = s0;
    } else {
        s0 = 1e-3 * (y + 1);
        s1 = 1e-3;
        s2 = 0;
    }
    fs0m.v = s0;
    if (!(is_active && fs0m.has_gradient)) return;
    // d(arg_a)/d(pos_b) = -J[a][b]
    const double q0 = -(pv.J[0][0] * g[0] + pv.J[1][0] * g[1]);  // dy/dpos1
    const double q1 = -(pv.J[0][1] * g[0] + pv.J[1][1] * g[1]);  // dy/dpos2
    fs0m.d[0] = s1 * q0;
    fs0m.d[1] = s1 * q1;
    if (!fs0m.has_hessian) return;
    // d2y/dpos_a dpos_b = sum_{cd} J[c][a] Hxy[c][d] J[d][b]
    const double Hm[2][2] = {{Hh[0], Hh[1]}, {Hh[1], Hh[2]}};
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            double yy = 0;
            for (int c = 0; c < 2; ++c)
                for (int d = 0; d < 2; ++d) yy += pv.J[c][a] * Hm[c][d] * pv.J[d][b];
            const double qa = a == 0 ? q0 : q1, qb = b == 0 ? q0 : q1;
            fs0m.H(a, b) = s2 * qa * qb + s1 * yy;
        }
}

// fsm_util.jl:255-346.  fs1m: SF(6,1); gal_ids: pos 0,1; frac_dev 2; shape 3,4,5
void accum_galaxy_pos(SF& fs1m, BvnDerivs& bd, const GalaxyCacheComponent& gcc, const double x[2],
                      const double J[2][2], bool is_active) {
    eval_bvn_pdf(bd, gcc.bmc, x);
    const double f = bd.f_pre * gcc.gal_frac_dev_i;
    fs1m.v += f;
    if (!(fs1m.has_gradient && is_active)) return;
    get_bvn_derivs(bd, gcc.bmc, fs1m.has_gradient, fs1m.has_hessian);
    transform_bvn_derivs(bd, gcc.sig_sf, J, fs1m.has_hessian);
    for (int u = 0; u < 2; ++u) fs1m.d[u] += f * bd.bvn_u_d[u];
    for (int g = 0; g < 3; ++g) fs1m.d[GAL_SHAPE_ALIGN[g]] += f * bd.bvn_s_d[g];
    fs1m.d[2] += gcc.gal_frac_dev_dir * bd.f_pre;
    if (!fs1m.has_hessian) return;
    for (int s1 = 0; s1 < 3; ++s1)
        for (int s2 = 0; s2 < 3; ++s2)
            fs1m.H(GAL_SHAPE_ALIGN[s1], GAL_SHAPE_ALIGN[s2]) += f * (bd.bvn_ss_h[s1][s2] + bd.bvn_s_d[s1] * bd.bvn_s_d[s2]);
    for (int u1 = 0; u1 < 2; ++u1)
        for (int u2 = 0; u2 < 2; ++u2) fs1m.H(u1, u2) += f * (bd.bvn_uu_h[u1][u2] + bd.bvn_u_d[u1] * bd.bvn_u_d[u2]);
    for (int u = 0; u < 2; ++u)
        for (int s = 0; s < 3; ++s) {
            const int si = GAL_SHAPE_ALIGN[s];
            fs1m.H(u, si) += f * (bd.bvn_us_h[u][s] + bd.bvn_u_d[u] * bd.bvn_s_d[s]);
            fs1m.H(si, u) = fs1m.H(u, si);
        }
    const int devi = 2;
    for (int u = 0; u < 2; ++u) {
        fs1m.H(u, devi) += bd.f_pre * gcc.gal_frac_dev_dir * bd.bvn_u_d[u];
        fs1m.H(devi, u) = fs1m.H(u, devi);
    }
    for (int s = 0; s < 3; ++s) {
        const int si = GAL_SHAPE_ALIGN[s];
        fs1m.H(si, devi) += bd.f_pre * gcc.gal_frac_dev_dir * bd.bvn_s_d[s];
        fs1m.H(devi, si) = fs1m.H(si, devi);
    }
}

// ---- deterministic_vi/elbo_args.jl:5-113 ---------------------------------------------
struct HessianSubmatrices {
    double u_u[2][2];
    std::vector<double> shape_shape;  // p x p col-major
    int p;
};
struct ElboVars {
    SF fs0m, fs1m, E_G_s, E_G2_s, var_G_s, E_G, var_G, elbo_log_term, elbo;
    HessianSubmatrices E_G_s_hsub[NT], E_G2_s_hsub[NT];
    double combine_grad[2];
    double combine_hess[2][2];
    int64_t active_pixel_counter = 0, inactive_pixel_counter = 0;
    BvnDerivs bd;
    ElboVars(int Sa, bool g, bool h)
        : fs0m(2, 1, g, h), fs1m(6, 1, g, h), E_G_s(P, 1, g, h), E_G2_s(P, 1, g, h), var_G_s(P, 1, g, h),
          E_G(P, Sa, g, h), var_G(P, Sa, g, h), elbo_log_term(P, Sa, g, h), elbo(P, Sa, g, h) {
        for (int i = 0; i < NT; ++i) {
            const int p = align().n_shape[i];
            E_G_s_hsub[i].p = E_G2_s_hsub[i].p = p;
            E_G_s_hsub[i].shape_shape.assign((size_t)p * p, 0.0);
            E_G2_s_hsub[i].shape_shape.assign((size_t)p * p, 0.0);
            std::memset(E_G_s_hsub[i].u_u, 0, sizeof E_G_s_hsub[i].u_u);
            std::memset(E_G2_s_hsub[i].u_u, 0, sizeof E_G2_s_hsub[i].u_u);
        }
        std::memset(&bd, 0, sizeof bd);
        combine_grad[0] = combine_grad[1] = 0;
        std::memset(combine_hess, 0, sizeof combine_hess);
    }
};
// elbo_args.jl:116-138
void zero(ElboVars& ev) {
    zero(ev.fs0m);
    zero(ev.fs1m);
    zero(ev.E_G_s);
    zero(ev.E_G2_s);
    zero(ev.var_G_s);
    for (int i = 0; i < NT; ++i) {
        std::memset(ev.E_G_s_hsub[i].u_u, 0, sizeof ev.E_G_s_hsub[i].u_u);
        std::memset(ev.E_G2_s_hsub[i].u_u, 0, sizeof ev.E_G2_s_hsub[i].u_u);
        std::fill(ev.E_G_s_hsub[i].shape_shape.begin(), ev.E_G_s_hsub[i].shape_shape.end(), 0.0);
        std::fill(ev.E_G2_s_hsub[i].shape_shape.begin(), ev.E_G2_s_hsub[i].shape_shape.end(), 0.0);
    }
    zero(ev.E_G);
    zero(ev.var_G);
    ev.combine_grad[0] = ev.combine_grad[1] = 0;
    std::memset(ev.combine_hess, 0, sizeof ev.combine_hess);
    zero(ev.elbo_log_term);
    zero(ev.elbo);
}

// ---- the ElboArgs view -------------------------------------------------------------------
struct Ea {
    int S, Sa, N, psf_K;
    const celeste_image* images;               // N
    std::vector<const celeste_patch*> patches;  // S x N (s + n*S)
    std::vector<int> active_sources;            // 0-based local indices
    const double* vp;                           // 44 x S
    const celeste_patch* patch(int s, int n) const { return patches[(size_t)s + (size_t)n * S]; }
    const double* vs(int s) const { return vp + (size_t)P * s; }
    int find_active(int s) const {
        for (size_t k = 0; k < active_sources.size(); ++k)
            if (active_sources[k] == s) return (int)k;
        return -1;
    }
};

// ---- elbo_objective.jl:17-233 --------------------------------------------------------------
void calculate_G_s(const Ea& ea, ElboVars& ev, const SourceBrightness& sb, int b, int s, bool is_active) {
    SF& E_G_s = ev.E_G_s;
    SF& E_G2_s = ev.E_G2_s;
    SF& var_G_s = ev.var_G_s;
    const Align& al = align();
    if (is_active) {
        zero(E_G_s);
        zero(E_G2_s);
        zero(var_G_s);
    } else {
        E_G_s.v = 0;
        E_G2_s.v = 0;
        var_G_s.v = 0;
    }
    const double* vps = ea.vs(s);
    for (int i = 0; i < NT; ++i) {
        const SF& fsm_i = (i == 0) ? ev.fs0m : ev.fs1m;
        const double a_i = vps[id_is_star(i)];
        const SF& El = sb.E_l_a[b][i];
        const SF& Ell = sb.E_ll_a[b][i];
        const double fsm_i_v = fsm_i.v, El_v = El.v, Ell_v = Ell.v;
        const double lf = El_v * fsm_i_v;
        const double llff = Ell_v * (fsm_i_v * fsm_i_v);
        E_G_s.v += a_i * lf;
        E_G2_s.v += a_i * llff;
        if (!(is_active && ev.elbo.has_gradient)) continue;

        E_G_s.d[id_is_star(i)] += lf;
        E_G2_s.d[id_is_star(i)] += llff;
        const int* p0_shape = al.shape[i];
        const int n_shape = al.n_shape[i];
        const int* p0_bright = al.bright[i];
        // u_ind = star_ids.pos / gal_ids.pos: both (0,1)
        const double tmp1 = El_v * a_i;
        const double tmp2 = Ell_v * 2 * fsm_i_v * a_i;
        for (int k = 0; k < n_shape; ++k) {
            E_G_s.d[p0_shape[k]] += tmp1 * fsm_i.d[k];
            E_G2_s.d[p0_shape[k]] += tmp2 * fsm_i.d[k];
        }
        for (int k = 0; k < BP; ++k) {
            E_G_s.d[p0_bright[k]] = a_i * fsm_i_v * El.d[k];
            E_G2_s.d[p0_bright[k]] = a_i * (fsm_i_v * fsm_i_v) * Ell.d[k];
        }
        if (!ev.elbo.has_hessian) continue;

        HessianSubmatrices& hs = ev.E_G_s_hsub[i];
        HessianSubmatrices& hs2 = ev.E_G2_s_hsub[i];
        // (bright, bright) :103-108
        for (int k1 = 0; k1 < BP; ++k1)
            for (int k2 = 0; k2 < BP; ++k2) {
                E_G_s.H(p0_bright[k1], p0_bright[k2]) = a_i * El.H(k1, k2) * fsm_i_v;
                E_G2_s.H(p0_bright[k1], p0_bright[k2]) = (fsm_i_v * fsm_i_v) * a_i * Ell.H(k1, k2);
            }
        // (shape, shape) :111-119
        const int p = hs.p;
        for (int i1 = 0; i1 < p; ++i1)
            for (int i2 = 0; i2 < p; ++i2) {
                hs.shape_shape[i1 + (size_t)i2 * p] = a_i * El_v * fsm_i.H(i1, i2);
                hs2.shape_shape[i1 + (size_t)i2 * p] =
                    2 * a_i * Ell_v * (fsm_i_v * fsm_i.H(i1, i2) + fsm_i.d[i1] * fsm_i.d[i2]);
            }
        // :123-128
        for (int k1 = 0; k1 < n_shape; ++k1)
            for (int k2 = 0; k2 < n_shape; ++k2) {
                E_G_s.H(p0_shape[k1], p0_shape[k2]) = a_i * El_v * fsm_i.H(k1, k2);
                E_G2_s.H(p0_shape[k1], p0_shape[k2]) = hs2.shape_shape[k1 + (size_t)k2 * p];
            }
        // :132-137
        for (int u1 = 0; u1 < 2; ++u1)
            for (int u2 = 0; u2 < 2; ++u2) {
                hs.u_u[u1][u2] = hs.shape_shape[u1 + (size_t)u2 * p];
                hs2.u_u[u1][u2] = hs2.shape_shape[u1 + (size_t)u2 * p];
            }
        // (a, bright) :144-153
        for (int k = 0; k < BP; ++k) {
            E_G_s.H(p0_bright[k], id_is_star(i)) = fsm_i_v * El.d[k];
            E_G2_s.H(p0_bright[k], id_is_star(i)) = (fsm_i_v * fsm_i_v) * Ell.d[k];
            E_G_s.H(id_is_star(i), p0_bright[k]) = E_G_s.H(p0_bright[k], id_is_star(i));
            E_G2_s.H(id_is_star(i), p0_bright[k]) = E_G2_s.H(p0_bright[k], id_is_star(i));
        }
        // (a, shape) :156-165
        for (int k = 0; k < n_shape; ++k) {
            E_G_s.H(p0_shape[k], id_is_star(i)) = El_v * fsm_i.d[k];
            E_G2_s.H(p0_shape[k], id_is_star(i)) = Ell_v * 2 * fsm_i_v * fsm_i.d[k];
            E_G_s.H(id_is_star(i), p0_shape[k]) = E_G_s.H(p0_shape[k], id_is_star(i));
            E_G2_s.H(id_is_star(i), p0_shape[k]) = E_G2_s.H(p0_shape[k], id_is_star(i));
        }
        // (bright, shape) :167-177
        for (int kb = 0; kb < BP; ++kb)
            for (int ks = 0; ks < n_shape; ++ks) {
                E_G_s.H(p0_bright[kb], p0_shape[ks]) = a_i * El.d[kb] * fsm_i.d[ks];
                E_G2_s.H(p0_bright[kb], p0_shape[ks]) = 2 * a_i * Ell.d[kb] * fsm_i_v * fsm_i.d[ks];
                E_G_s.H(p0_shape[ks], p0_bright[kb]) = E_G_s.H(p0_bright[kb], p0_shape[ks]);
                E_G2_s.H(p0_shape[ks], p0_bright[kb]) = E_G2_s.H(p0_bright[kb], p0_shape[ks]);
            }
    }
    // :180-199 (pos x pos summed over the two types)
    if (ev.elbo.has_hessian) {
        const int ipos[2] = {ID_POS0, ID_POS1};
        for (int u1 = 0; u1 < 2; ++u1)
            for (int u2 = 0; u2 < 2; ++u2) {
                E_G_s.H(ipos[u1], ipos[u2]) = ev.E_G_s_hsub[0].u_u[u1][u2] + ev.E_G_s_hsub[1].u_u[u1][u2];
                E_G2_s.H(ipos[u1], ipos[u2]) = ev.E_G2_s_hsub[0].u_u[u1][u2] + ev.E_G2_s_hsub[1].u_u[u1][u2];
            }
    }
    // :204
    var_G_s.v = E_G2_s.v - (E_G_s.v * E_G_s.v);
    if (!(is_active && ev.elbo.has_gradient)) return;
    // :215-217
    for (int k = 0; k < P; ++k) var_G_s.d[k] = E_G2_s.d[k] - 2 * E_G_s.v * E_G_s.d[k];
    if (!ev.elbo.has_hessian) return;
    // :226-232
    for (int i2 = 0; i2 < P; ++i2)
        for (int i1 = 0; i1 <= i2; ++i1) {
            var_G_s.H(i1, i2) = E_G2_s.H(i1, i2) - 2 * (E_G_s.v * E_G_s.H(i1, i2) + E_G_s.d[i1] * E_G_s.d[i2]);
            var_G_s.H(i2, i1) = var_G_s.H(i1, i2);
        }
}

// :240-259
void accumulate_source_pixel_brightness(const Ea& ea, ElboVars& ev, const SourceBrightness& sb, int b, int s,
                                        bool is_active) {
    calculate_G_s(ea, ev, sb, b, s, is_active);
    if (is_active) {
        const int sa = ea.find_active(s);
        add_sources_sf(ev.E_G, ev.E_G_s, sa);
        add_sources_sf(ev.var_G, ev.var_G_s, sa);
    } else {
        ev.E_G.v += ev.E_G_s.v;
        ev.var_G.v += ev.var_G_s.v;
    }
}

// :274-327.  log_iota = Float64(log(iota::Float32)) (:292)
void add_elbo_log_term(ElboVars& ev, double x_nbm, double log_iota) {
    const double E_G_v = ev.E_G.v, var_G_v = ev.var_G.v;
    const double log_term_value = std::log(E_G_v) - var_G_v / (2.0 * (E_G_v * E_G_v));
    ev.elbo.v += x_nbm * (log_iota + log_term_value);
    if (!ev.elbo.has_gradient) return;
    ev.combine_grad[0] = -0.5 / (E_G_v * E_G_v);
    ev.combine_grad[1] = 1 / E_G_v + var_G_v / (E_G_v * E_G_v * E_G_v);
    if (ev.elbo.has_hessian) {
        ev.combine_hess[0][0] = 0.0;
        ev.combine_hess[0][1] = ev.combine_hess[1][0] = 1 / (E_G_v * E_G_v * E_G_v);
        ev.combine_hess[1][1] = -(1 / (E_G_v * E_G_v) + 3 * var_G_v / (E_G_v * E_G_v * E_G_v * E_G_v));
    }
    combine_sfs(ev.var_G, ev.E_G, ev.elbo_log_term, log_term_value, ev.combine_grad, ev.combine_hess);
    for (size_t k = 0; k < ev.elbo.d.size(); ++k) ev.elbo.d[k] += x_nbm * ev.elbo_log_term.d[k];
    if (ev.elbo.has_hessian)
        for (size_t k = 0; k < ev.elbo.h.size(); ++k) ev.elbo.h[k] += x_nbm * ev.elbo_log_term.h[k];
}

struct Mixtures {  // BvnBundle.gal_mcs fsm_util.jl:68-77: [k][j][i][s]
    std::vector<GalaxyCacheComponent> gal;
    int K, S;
    GalaxyCacheComponent& at(int k, int j, int i, int s) { return gal[k + (size_t)K * (j + 8 * (i + 2 * (size_t)s))]; }
};

// fsm_util.jl:111-169 (star_mcs is dead code on this path; not built)
void load_bvn_mixtures(Mixtures& mx, const Ea& ea, int n, bool calc_grad, bool calc_hess) {
    const Prototypes& pr = prototypes();
    for (int s = 0; s < ea.S; ++s) {
        const celeste_patch* p = ea.patch(s, n);
        const PatchView pv = view(p);
        const double* sp = ea.vs(s);
        double m_pos[2];
        linear_world_to_pix(pv, sp + ID_POS0, m_pos);
        const bool active = ea.find_active(s) >= 0;
        for (int i = 0; i < 2; ++i) {
            const double dir = (i == 0) ? 1. : -1.;
            const double frac_i = (i == 0) ? sp[ID_FRAC_DEV] : 1. - sp[ID_FRAC_DEV];
            for (int j = 0; j < pr.n[i]; ++j)
                for (int k = 0; k < ea.psf_K; ++k)
                    mx.at(k, j, i, s) = make_gcc(dir, frac_i, pr.eta[i][j], pr.nu[i][j], p->psf + 7 * k, m_pos,
                                                 sp[ID_AXIS_RATIO], sp[ID_ANGLE], sp[ID_RADIUS], calc_grad && active,
                                                 calc_hess);
        }
    }
}

// fsm_util.jl:194-219
void populate_gal_fsm(SF& fs1m, BvnDerivs& bd, int s, int h, int w, bool is_active, const double J[2][2],
                      Mixtures& mx) {
    zero(fs1m);
    const double x[2] = {(double)h, (double)w};
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 8; ++j)
            if (i == 0 || j < 6)
                for (int k = 0; k < mx.K; ++k) accum_galaxy_pos(fs1m, bd, mx.at(k, j, i, s), x, J, is_active);
}

inline float img_px(const celeste_image& im, int h, int w) { return im.pixels[(size_t)(h - 1) + (size_t)(w - 1) * im.H]; }
inline float img_sky(const celeste_image& im, int h, int w) { return im.sky[(size_t)(h - 1) + (size_t)(w - 1) * im.H]; }
inline double img_log_iota(const celeste_image& im, int h) {
    return im.log_iota ? im.log_iota[h - 1] : (double)std::log(im.nelec_per_nmgy[h - 1]);  // Float32 log
}

// elbo_objective.jl:330-392 (h, w 1-based image coordinates)
void add_pixel_term(const Ea& ea, int n, int h, int w, Mixtures& mx, const std::vector<SourceBrightness>& sbs,
                    ElboVars& ev) {
    const celeste_image& img = ea.images[n];
    zero(ev.E_G);
    zero(ev.var_G);
    for (int s = 0; s < ea.S; ++s) {
        const celeste_patch* p = ea.patch(s, n);
        const int64_t h2 = h - p->bitmap_offset[0];
        const int64_t w2 = w - p->bitmap_offset[1];
        const int H2 = p->H2, W2 = p->W2;
        // NB the strict `w2 < W2` (:349)
        if (1 <= h2 && h2 <= H2 && 1 <= w2 && w2 < W2 && p->active_pixel_bitmap[(h2 - 1) + (size_t)(w2 - 1) * H2]) {
            const bool is_active = ea.find_active(s) >= 0;
            if (is_active)
                ev.active_pixel_counter += 1;
            else
                ev.inactive_pixel_counter += 1;
            const PatchView pv = view(p);
            star_light_density(ev.fs0m, pv, h, w, ea.vs(s) + ID_POS0, is_active);
            populate_gal_fsm(ev.fs1m, ev.bd, s, h, w, is_active, pv.J, mx);
            accumulate_source_pixel_brightness(ea, ev, sbs[s], img.band - 1, s, is_active);
        }
    }
    ev.E_G.v += img_sky(img, h, w);  // :374
    const float px = img_px(img, h, w);
    const float iota = img.nelec_per_nmgy[h - 1];
    add_elbo_log_term(ev, (double)px, img_log_iota(img, h));
    add_scaled_sfs(ev.elbo, ev.E_G, -(double)iota);  // :383-385
    ev.elbo.v -= std::lgamma((double)px + 1.0);      // :391
}

// source_brightness.jl:213-229
void load_source_brightnesses(const Ea& ea, std::vector<SourceBrightness>& sbs) {
    sbs.resize(ea.S);
    for (int s = 0; s < ea.S; ++s) {
        const bool act = ea.find_active(s) >= 0;
        source_brightness(ea.vs(s), act, act, sbs[s]);
    }
}

bool all_finite(const SF& sf) {  // elbo_args.jl:145-149
    if (!std::isfinite(sf.v)) return false;
    for (double x : sf.d)
        if (!std::isfinite(x)) return false;
    for (double x : sf.h)
        if (!std::isfinite(x)) return false;
    return true;
}

// elbo_objective.jl:400-474
void elbo_likelihood(const Ea& ea, ElboVars& ev) {
    zero(ev);
    ev.active_pixel_counter = ev.inactive_pixel_counter = 0;  // fresh counters per call (scratch is per call here)
    std::vector<SourceBrightness> sbs;
    load_source_brightnesses(ea, sbs);
    Mixtures mx;
    mx.K = ea.psf_K;
    mx.S = ea.S;
    mx.gal.resize((size_t)ea.psf_K * 8 * 2 * ea.S);
    std::vector<uint8_t> already_visited;
    for (int n = 0; n < ea.N; ++n) {
        const celeste_image& img = ea.images[n];
        load_bvn_mixtures(mx, ea, n, ev.elbo.has_gradient, ev.elbo.has_hessian);
        const bool dedupe = ea.active_sources.size() != 1;
        if (dedupe) already_visited.assign((size_t)img.H * img.W, 0);
        for (int s : ea.active_sources) {
            const celeste_patch* p = ea.patch(s, n);
            const int H2 = p->H2, W2 = p->W2;
            for (int w2 = 1; w2 <= W2; ++w2)
                for (int h2 = 1; h2 <= H2; ++h2) {
                    const int h = (int)(p->bitmap_offset[0] + h2);
                    const int w = (int)(p->bitmap_offset[1] + w2);
                    if (!p->active_pixel_bitmap[(h2 - 1) + (size_t)(w2 - 1) * H2]) continue;
                    if (dedupe) {
                        uint8_t& av = already_visited[(size_t)(h - 1) + (size_t)(w - 1) * img.H];
                        if (av) continue;
                        av = 1;
                    }
                    if (std::isnan(img_px(img, h, w))) continue;
                    add_pixel_term(ea, n, h, w, mx, sbs, ev);
                }
        }
    }
}

struct TaskOut {
    double* v;
    double* d;
    double* h;
    int64_t* counters;
    int32_t* flag;
};

void run_task(const celeste_image* imgs, int N, const celeste_patch* patches, int S_tot, int S, const int32_t* src_ids,
              int Sa, const int32_t* act_idx, const double* vp, int mode, TaskOut out) {
    Ea ea;
    ea.S = S;
    ea.Sa = Sa;
    ea.N = N;
    ea.images = imgs;
    ea.vp = vp;
    ea.patches.resize((size_t)S * N);
    for (int n = 0; n < N; ++n)
        for (int s = 0; s < S; ++s) ea.patches[s + (size_t)n * S] = &patches[(size_t)(src_ids[s] - 1) + (size_t)n * S_tot];
    ea.psf_K = S > 0 && N > 0 ? ea.patches[0]->K : 2;
    for (int k = 0; k < Sa; ++k) ea.active_sources.push_back(act_idx[k] - 1);
    ElboVars ev(Sa, mode >= 1, mode >= 2);
    elbo_likelihood(ea, ev);
    *out.v = ev.elbo.v;
    if (mode >= 1 && out.d) std::memcpy(out.d, ev.elbo.d.data(), ev.elbo.d.size() * sizeof(double));
    if (mode >= 2 && out.h) std::memcpy(out.h, ev.elbo.h.data(), ev.elbo.h.size() * sizeof(double));
    if (out.counters) {
        out.counters[0] = ev.active_pixel_counter;
        out.counters[1] = ev.inactive_pixel_counter;
    }
    if (out.flag) *out.flag = all_finite(ev.elbo) ? 0 : CELESTE_FLAG_NONFINITE;
}

}  // namespace

extern "C" {

// Same argument meaning as celeste_elbo_batch (include/celeste_cuda.h), with the
// images / patch matrix passed directly (host memory) and a thread count.  Worker
// threads pull the next task index from a shared counter exactly like
// one_node_single_infer (ParallelRun.jl:553-597), one scratch object per evaluation.
int oracle_elbo_batch(int32_t N, const celeste_image* imgs, int32_t S_tot, const celeste_patch* patches,
                      int32_t n_tasks, const int32_t* task_ptr, const int32_t* source_ids, const int32_t* active_ptr,
                      const int32_t* active_idx, const double* vp, int32_t mode, double* v, double* d, double* h,
                      int64_t* counters, int32_t* flags, int32_t n_threads) {
    std::vector<size_t> hptr(n_tasks + 1, 0);
    for (int t = 0; t < n_tasks; ++t) {
        const size_t pp = (size_t)P * (active_ptr[t + 1] - active_ptr[t]);
        hptr[t + 1] = hptr[t] + pp * pp;
    }
    std::atomic<int> next{0};
    auto worker = [&]() {
        for (;;) {
            const int t = next.fetch_add(1);
            if (t >= n_tasks) break;
            TaskOut o;
            o.v = v + t;
            o.d = d ? d + (size_t)P * active_ptr[t] : nullptr;
            o.h = h ? h + hptr[t] : nullptr;
            o.counters = counters ? counters + 2 * (size_t)t : nullptr;
            o.flag = flags ? flags + t : nullptr;
            run_task(imgs, N, patches, S_tot, task_ptr[t + 1] - task_ptr[t], source_ids + task_ptr[t],
                     active_ptr[t + 1] - active_ptr[t], active_idx + active_ptr[t], vp + (size_t)P * task_ptr[t], mode, o);
        }
    };
    if (n_threads <= 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < n_threads; ++i) th.emplace_back(worker);
        for (auto& x : th) x.join();
    }
    return 0;
}

// fill_celeste_expectation! (bin/write_celeste_expectation.jl:111-156): ElboArgs(images, patches, [1]) with
// value-only scratch, then for EVERY pixel of every image add_pixel_term! and `pixels[h, w] += E_G.v - sky[h, w]`.
// out[n] (H x W doubles, column-major) receives that increment.  Worker threads split each image by columns.
int oracle_render_expectation(int32_t N, const celeste_image* imgs, int32_t S_tot, const celeste_patch* patches,
                              int32_t S, const int32_t* source_ids, const double* vp, double* const* out,
                              int32_t n_threads) {
    Ea ea;
    ea.S = S;
    ea.Sa = S > 0 ? 1 : 0;
    ea.N = N;
    ea.images = imgs;
    ea.vp = vp;
    ea.patches.resize((size_t)S * N);
    for (int n = 0; n < N; ++n)
        for (int s = 0; s < S; ++s) ea.patches[s + (size_t)n * S] = &patches[(size_t)(source_ids[s] - 1) + (size_t)n * S_tot];
    ea.psf_K = S > 0 && N > 0 ? ea.patches[0]->K : 2;
    if (S > 0) ea.active_sources.push_back(0);                     // active_sources = [1]
    std::vector<SourceBrightness> sbs;
    load_source_brightnesses(ea, sbs);                             // derivative flags do not change the values
    for (int n = 0; n < N; ++n) {
        const celeste_image& img = imgs[n];
        Mixtures mx;
        mx.K = ea.psf_K;
        mx.S = S;
        mx.gal.resize((size_t)ea.psf_K * 8 * 2 * S);
        load_bvn_mixtures(mx, ea, n, false, false);
        std::atomic<int> next{1};
        auto worker = [&]() {
            ElboVars ev(std::max(ea.Sa, 1), false, false);
            for (;;) {
                const int w = next.fetch_add(1);
                if (w > img.W) break;
                for (int h = 1; h <= img.H; ++h) {
                    // the E_G part of add_pixel_term! (:340-374)
                    zero(ev.E_G);
                    zero(ev.var_G);
                    for (int s = 0; s < S; ++s) {
                        const celeste_patch* p = ea.patch(s, n);
                        const int64_t h2 = h - p->bitmap_offset[0], w2 = w - p->bitmap_offset[1];
                        if (1 <= h2 && h2 <= p->H2 && 1 <= w2 && w2 < p->W2 &&
                            p->active_pixel_bitmap[(h2 - 1) + (size_t)(w2 - 1) * p->H2]) {
                            const bool is_active = ea.find_active(s) >= 0;
                            const PatchView pv = view(p);
                            star_light_density(ev.fs0m, pv, h, w, ea.vs(s) + ID_POS0, is_active);
                            Mixtures& mref = mx;
                            populate_gal_fsm(ev.fs1m, ev.bd, s, h, w, is_active, pv.J, mref);
                            accumulate_source_pixel_brightness(ea, ev, sbs[s], img.band - 1, s, is_active);
                        }
                    }
                    const double sky = (double)img_sky(img, h, w);
                    ev.E_G.v += sky;
                    out[n][(size_t)(h - 1) + (size_t)(w - 1) * img.H] = ev.E_G.v - sky;
                }
            }
        };
        if (n_threads <= 1) {
            worker();
        } else {
            std::vector<std::thread> th;
            for (int i = 0; i < n_threads; ++i) th.emplace_back(worker);
            for (auto& x : th) x.join();
        }
    }
    return 0;
}

// get_bvn_cov closed form (test/test_elbo.jl:45-61): out = (S11, S12, S22)
void oracle_get_bvn_cov(double ab, double angle, double scale, double* out) {
    double c[2][2];
    get_bvn_cov(ab, angle, scale, c);
    out[0] = c[0][0];
    out[1] = c[0][1];
    out[2] = c[1][1];
}

// Spline evaluation alone (value, gradient, Hessian) for the interpolant self-check.
void oracle_spline_eval(const double* coefs, int32_t n1, int32_t n2, double x, double y, double* out6) {
    double v, g[2], H[3];
    spline_eval(coefs, n1, n2, x, y, v, g, H);
    out6[0] = v;
    out6[1] = g[0];
    out6[2] = g[1];
    out6[3] = H[0];
    out6[4] = H[1];
    out6[5] = H[2];
}

// Galaxy prototypes (light_source_model.jl:45-72): eta[2][8], nu[2][8]
void oracle_galaxy_prototypes(double* eta16, double* nu16) {
    const Prototypes& p = prototypes();
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 8; ++j) {
            eta16[i * 8 + j] = p.eta[i][j];
            nu16[i * 8 + j] = p.nu[i][j];
        }
}

// calculate_G_s! probe for the "overwrites" property test (test/test_elbo.jl:13-42):
// runs elbo_likelihood once (to leave non-trivial scratch), then the given sequence
// of calculate_G_s! calls (s, b 1-based) and returns E_G_s/E_G2_s/var_G_s (v, d[44], h[44*44]) x 3.
int oracle_calculate_G_s_probe(int32_t N, const celeste_image* imgs, int32_t S, const celeste_patch* patches,
                               int32_t Sa, const int32_t* active_idx, const double* vp, int32_t n_calls,
                               const int32_t* call_s, const int32_t* call_b, double* out) {
    Ea ea;
    ea.S = S;
    ea.Sa = Sa;
    ea.N = N;
    ea.images = imgs;
    ea.vp = vp;
    ea.patches.resize((size_t)S * N);
    for (int n = 0; n < N; ++n)
        for (int s = 0; s < S; ++s) ea.patches[s + (size_t)n * S] = &patches[(size_t)s + (size_t)n * S];
    ea.psf_K = ea.patches[0]->K;
    for (int k = 0; k < Sa; ++k) ea.active_sources.push_back(active_idx[k] - 1);
    ElboVars ev(Sa, true, true);
    elbo_likelihood(ea, ev);
    std::vector<SourceBrightness> sbs;
    load_source_brightnesses(ea, sbs);
    for (int c = 0; c < n_calls; ++c) calculate_G_s(ea, ev, sbs[call_s[c] - 1], call_b[c] - 1, call_s[c] - 1, true);
    const SF* sfs[3] = {&ev.E_G_s, &ev.E_G2_s, &ev.var_G_s};
    size_t o = 0;
    for (const SF* sf : sfs) {
        out[o++] = sf->v;
        for (double x : sf->d) out[o++] = x;
        for (double x : sf->h) out[o++] = x;
    }
    return 0;
}

}  // extern "C"
