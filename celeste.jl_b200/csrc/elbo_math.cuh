// elbo_math.cuh -- per-pixel arithmetic of the ELBO hot path, written B200-first.
//
// The reference evaluates, for every (pixel, source), 28 Gaussian components and pushes
// each one through a dense chain rule into a 44 x 44 SensitiveFloat
// (BivariateNormals.jl:208-571, fsm_util.jl:255-346, elbo_objective.jl:17-327,
// SensitiveFloats.jl:99-250).  This file computes the same function, reorganised so the
// per-pixel work is a short FP64 FMA stream with no memory traffic:
//
//  * all 14K components of a source share one shape Jacobian up to the scalar nuBar
//    (GalaxySigmaDerivs returns j*nuBar, t*nuBar: BivariateNormals.jl:396) and one
//    position Jacobian -J (transform_bvn_ux_derivs!:414), so derivatives are accumulated
//    in the RAW coordinates y = (x1, x2, Sigma11, Sigma12, Sigma22, theta) with weights
//    w, w*nuBar, w*nuBar^2 and transformed to (pos, gal_frac_dev, axis_ratio, angle,
//    radius) ONCE per (source, image) in the epilogue -- not once per component;
//  * the pixel term depends on the 28 live parameters only through the raw coordinates
//    and four per-band scalars c = (a1 E_l1, a2 E_l2, a1 E_ll1, a2 E_ll2)
//    (calculate_G_s!, elbo_objective.jl:62-66), so each thread accumulates the
//    value / gradient / Hessian of the pixel term in (c, y) space: 1 + 10 + 55 numbers.
//
// Mathematically identical to the reference; differs by floating-point reassociation
// only (parity tolerance 1e-8).
//
// Functions are __host__ __device__ so that tests/host_emul can compile this exact
// arithmetic with g++ and compare it with the oracle on a machine without a GPU.  The
// shipped library only ever calls them from kernels.
#ifndef CELESTE_ELBO_MATH_CUH
#define CELESTE_ELBO_MATH_CUH

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CEL_HD __host__ __device__ __forceinline__
#else
#define CEL_HD inline
#endif

namespace celeste {

constexpr int NPARAM = 44;     // length(CanonicalParams), param_set.jl:107
constexpr int NLIVE = 28;      // canonical ids 1..28 receive likelihood derivatives (ids.k never does)
constexpr int NPROTO = 14;     // 8 dev + 6 exp prototype components, light_source_model.jl:45-72
constexpr int NPROTO_DEV = 8;
constexpr int MAX_K = 4;       // PSF components supported per patch (reference default psf_K = 2)
constexpr int COMP_STRIDE = 8; // doubles per component record: mu1 mu2 L11 L12 L22 z L11/2 L22/2
constexpr int MAX_COMPS = NPROTO * MAX_K;

// accumulator layout in (c, y) space.  y order: x1 x2 S11 S12 S22 theta; c order: A1 A2 B1 B2
constexpr int ACC_VAL = 0, ACC_CNT_ACTIVE = 1, ACC_CNT_INACTIVE = 2;
constexpr int ACC_G = 3;        // 6: dL/dy
constexpr int ACC_C1 = 9;       // 4: dL/dc
constexpr int ACC_HH = 13;      // 21: d2L/dy dy, packed upper triangle row-major (k<=l)
constexpr int ACC_CC = 34;      // 10: d2L/dc dc, packed upper triangle
constexpr int ACC_CR = 44;      // 24: d2L/dc dy, [c][k]
constexpr int NACC_MODE0 = 3, NACC_MODE1 = 13, NACC_MODE2 = 68;
template <int MODE> struct NAcc { static constexpr int value = MODE == 0 ? NACC_MODE0 : (MODE == 1 ? NACC_MODE1 : NACC_MODE2); };
constexpr int HESS_PACKED_LEN = NLIVE * (NLIVE + 1) / 2;   // 406: upper triangle of the live 28 x 28 block (CELESTE_HESS_PACKED28)
CEL_HD constexpr int hess_packed_index(int r, int c) { return r * NLIVE - (r * (r - 1)) / 2 + (c - r); }   // r <= c < 28
CEL_HD constexpr int tri6(int k, int l) { return k * 6 - (k * (k - 1)) / 2 + (l - k); }   // k <= l < 6
CEL_HD constexpr int tri4(int k, int l) { return k * 4 - (k * (k - 1)) / 2 + (l - k); }   // k <= l < 4

// ---------------------------------------------------------------------------------------------
// exp(s*q) for s*q <= 0 (the only arguments this path produces: -q/2 of a Gaussian, and the negative
// branch of softpluslikeinv), computed as 2^y with y = q * (s * log2 e):
//   k = rint(y) by the 1.5*2^52 shift, f = y - k EXACTLY (no Cody-Waite constants needed), |f| <= 1/2,
//   2^f by a degree-10 near-minimax polynomial (Chebyshev interpolant; max relative error 6.7e-16
//   measured against 50-digit arithmetic, tools/fit_exp.py), Estrin order for instruction-level
//   parallelism, exponent spliced in with integer arithmetic.
// No slow path and no FP64 compare: the integer exponent is clamped at -1022 (the result is then
// ~1e-308 instead of a denormal/zero, far below anything the ELBO can resolve).  The rounding of y
// costs |y| * 1.1e-16 relative accuracy (1e-14 at q = 200, where the term is already e^-100).
// On the device the coefficients sit in constant memory so every DFMA takes its constant as a
// c[bank][offset] operand instead of two UMOVs.
#define CEL_EXP2_COEFS                                                                                          \
    {1.0, 0.69314718055994995, 0.24022650695910097, 0.055504108664447417, 0.0096181291076068709,               \
     0.0013333558230215262, 0.00015403530441765088, 1.5252657229551429e-05, 1.3215442570224649e-06,            \
     1.0208696429374313e-07, 7.0725894883963448e-09}
#if defined(__CUDACC__)
__constant__ double c_expc[11] = CEL_EXP2_COEFS;
#endif
CEL_HD double exp_scaled(double q, double s) {
#if defined(__CUDA_ARCH__)
    const double* C = c_expc;
#else
    static const double C[11] = CEL_EXP2_COEFS;
#endif
    const double SHIFT = 6755399441055744.0;   // 1.5 * 2^52
    const double y = q * (s * 1.4426950408889634074);
    const double kd = y + SHIFT;
    const double r = y - (kd - SHIFT);
    const double r2 = r * r;
    const double a0 = fma(C[1], r, C[0]);
    const double a1 = fma(C[3], r, C[2]);
    const double a2 = fma(C[5], r, C[4]);
    const double a3 = fma(C[7], r, C[6]);
    const double a4 = fma(C[9], r, C[8]);
    const double r4 = r2 * r2;
    const double b0 = fma(a1, r2, a0);
    const double b1 = fma(a3, r2, a2);
    const double b2 = fma(C[10], r2, a4);
    const double r8 = r4 * r4;
    const double p = fma(b2, r8, fma(b1, r4, b0));
#if defined(__CUDA_ARCH__)
    int k = __double2loint(kd);
    k = k < -1022 ? -1022 : k;
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
    long long bits, kb;
    memcpy(&kb, &kd, 8);
    int k = (int)(kb & 0xffffffffLL);
    k = k < -1022 ? -1022 : k;
    memcpy(&bits, &p, 8);
    bits += (long long)k << 52;
    double out;
    memcpy(&out, &bits, 8);
    return out;
#endif
}
CEL_HD double exp_nonpos(double x) { return exp_scaled(x, 1.0); }

// Table-driven variant for the mixture loops (28 calls per pixel-source): 2^y = 2^e * T[j] * 2^r with
// 8 y = 8 e + j + 8 r, T[j] = 2^(j/8) from an 8-entry table (64 bytes of shared memory: lanes with equal j
// broadcast, different j hit different banks), |r| <= 1/16 and a degree-6 near-minimax polynomial
// (max relative error 1.1e-15 against 50-digit arithmetic).  12 FP64 instructions instead of 17.
#define CEL_EXP2_TAB                                                                                                 \
    {1.0, 1.0905077326652577, 1.189207115002721, 1.2968395546510096, 1.4142135623730951, 1.5422108254079407,        \
     1.681792830507429, 1.8340080864093424}
#define CEL_EXP2_C6                                                                                                  \
    {1.0, 0.69314718056004476, 0.24022650695910933, 0.055504108461085513, 0.0096181290899762413,                    \
     0.0013334601056309483, 0.00015404434000046519}
#if defined(__CUDACC__)
__constant__ double c_exp6[7] = CEL_EXP2_C6;
__constant__ double c_exptab[8] = CEL_EXP2_TAB;
#endif
CEL_HD double exp_scaled_tab(double q, double s, const double* tab) {
#if defined(__CUDA_ARCH__)
    const double* C = c_exp6;
#else
    static const double C[7] = CEL_EXP2_C6;
#endif
    const double SHIFT = 6755399441055744.0;   // 1.5 * 2^52
    const double y = q * (s * 1.4426950408889634074);
    const double kd = fma(y, 8.0, SHIFT);       // k8 = rint(8 y)
    const double r = fma(kd - SHIFT, -0.125, y);
    const double r2 = r * r;
    const double a0 = fma(C[1], r, C[0]);
    const double a1 = fma(C[3], r, C[2]);
    const double a2 = fma(C[5], r, C[4]);
    const double p = fma(fma(fma(C[6], r2, a2), r2, a1), r2, a0);
#if defined(__CUDA_ARCH__)
    const int k8 = __double2loint(kd);
#else
    long long kb;
    memcpy(&kb, &kd, 8);
    const int k8 = (int)(kb & 0xffffffffLL);
#endif
    int e = k8 >> 3;
    const double pt = p * tab[k8 & 7];
    e = e < -1022 ? -1022 : e;
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(pt) + (e << 20), __double2loint(pt));
#else
    long long bits;
    memcpy(&bits, &pt, 8);
    bits += (long long)e << 52;
    double out;
    memcpy(&out, &bits, 8);
    return out;
#endif
}
#if !defined(__CUDACC__)
static const double h_exptab[8] = CEL_EXP2_TAB;
#endif

// ---------------------------------------------------------------------------------------------
// log(E) for the pixel term (add_elbo_log_term!, elbo_objective.jl:288-292), table-driven: E = 2^k m, m in [1, 2),
// i = top 7 mantissa bits, c_i = 1 + (i + 1/2)/128;  log E = k ln 2 - log(inv_i) + log1p(r),  r = fma(m, inv_i, -1)
// with inv_i = fl(1 / c_i) (the identity is exact for the STORED inv_i; logc_i = -log(inv_i) rounded once),
// |r| <= 2^-8, log1p by a degree-7 Taylor polynomial (|r|^8 / 8 < 1e-20).  About 20 instructions instead of the
// ~80 of the library routine (which divides); absolute error ~1e-16 (tools/fit_exp.py-style check against 50-digit
// arithmetic: <= 2e-14 relative to max(|log E|, 1e-3)).  Non-positive, subnormal, infinite and NaN arguments take
// the library routine, so the non-finite flag behaves as before.
// Table: 128 x (inv_i, logc_i).
#define CEL_LOG_TAB { \
    0.9961089494163424, 0.003898640415657309, \
    0.9884169884169884, 0.01165061721997525, \
    0.9808429118773946, 0.019342962843130987, \
    0.973384030418251, 0.026976587698202083, \
    0.9660377358490566, 0.03455238150665973, \
    0.9588014981273408, 0.042071213920687044, \
    0.9516728624535316, 0.049533935122276676, \
    0.9446494464944649, 0.05694137640013845, \
    0.9377289377289377, 0.06429435070539725, \
    0.9309090909090909, 0.07159365318700882, \
    0.924187725631769, 0.078840061707776, \
    0.9175627240143369, 0.08603433734180316, \
    0.9110320284697508, 0.09317722485418334, \
    0.9045936395759717, 0.10026945316367517, \
    0.8982456140350877, 0.10731173578908804, \
    0.89198606271777, 0.11430477128005863, \
    0.8858131487889274, 0.12124924363286965, \
    0.8797250859106529, 0.12814582269193006, \
    0.8737201365187713, 0.13499516453750482, \
    0.8677966101694915, 0.1417979118602574, \
    0.8619528619528619, 0.1485546943231372, \
    0.8561872909698997, 0.15526612891112396, \
    0.8504983388704319, 0.16193282026931324, \
    0.8448844884488449, 0.16855536102980664, \
    0.839344262295082, 0.17513433212784915, \
    0.8338762214983714, 0.18167030310763463, \
    0.8284789644012945, 0.18816383241818294, \
    0.8231511254019293, 0.19461546769967167, \
    0.8178913738019169, 0.2010257460605908, \
    0.8126984126984127, 0.2073951943460706, \
    0.807570977917981, 0.21372432939771818, \
    0.8025078369905956, 0.22001365830528213, \
    0.7975077881619937, 0.2262636786504534, \
    0.7925696594427245, 0.232474878743094, \
    0.7876923076923077, 0.238647737850175, \
    0.7828746177370031, 0.24478272641769092, \
    0.7781155015197568, 0.25088030628580943, \
    0.7734138972809668, 0.2569409308975004, \
    0.7687687687687688, 0.26296504550088134, \
    0.764179104477612, 0.26895308734550394, \
    0.7596439169139466, 0.2749054858727992, \
    0.7551622418879056, 0.2808226629008878, \
    0.750733137829912, 0.2867050328039543, \
    0.7463556851311953, 0.29255300268637746, \
    0.7420289855072464, 0.2983669725517973, \
    0.7377521613832853, 0.3041473354672968, \
    0.7335243553008596, 0.3098944777228647, \
    0.7293447293447294, 0.3156087789863033, \
    0.7252124645892352, 0.32129061245373425, \
    0.7211267605633803, 0.3269403449958533, \
    0.7170868347338936, 0.3325583373000766, \
    0.713091922005571, 0.3381449440087164, \
    0.7091412742382271, 0.34370051385331846, \
    0.7052341597796143, 0.3492253897852883, \
    0.7013698630136986, 0.354719909102929, \
    0.6975476839237057, 0.3601844035750078, \
    0.6937669376693767, 0.3656191995609647, \
    0.6900269541778976, 0.37102461812787263, \
    0.6863270777479893, 0.376400975164253, \
    0.6826666666666666, 0.3817485814908484, \
    0.6790450928381963, 0.3870677429684483, \
    0.6754617414248021, 0.3923587606028639, \
    0.6719160104986877, 0.3976219306471385, \
    0.6684073107049608, 0.4028575447010835, \
    0.6649350649350649, 0.4080658898082217, \
    0.661498708010336, 0.41324724855021927, \
    0.6580976863753213, 0.41840189913888387, \
    0.6547314578005116, 0.4235301155058032, \
    0.6513994910941476, 0.42863216738969867, \
    0.6481012658227848, 0.4337083204215594, \
    0.6448362720403022, 0.43875883620762796, \
    0.6416040100250626, 0.44378397241030104, \
    0.6384039900249376, 0.4487839828270067, \
    0.6352357320099256, 0.4537591174671205, \
    0.6320987654320988, 0.4587096226269767, \
    0.628992628992629, 0.46363574096303256, \
    0.6259168704156479, 0.46853771156323926, \
    0.6228710462287105, 0.4734157700166721, \
    0.6198547215496368, 0.47827014848147026, \
    0.6168674698795181, 0.48310107575113576, \
    0.6139088729016786, 0.48790877731923904, \
    0.6109785202863962, 0.4926934754425752, \
    0.6080760095011877, 0.4974553892028189, \
    0.6052009456264775, 0.5021947345667155, \
    0.6023529411764705, 0.5069117244448544, \
    0.5995316159250585, 0.5116065687490621, \
    0.5967365967365967, 0.5162794744484545, \
    0.5939675174013921, 0.5209306456241853, \
    0.5912240184757506, 0.5255602835229274, \
    0.5885057471264368, 0.5301685866091216, \
    0.585812356979405, 0.5347557506160276, \
    0.5831435079726651, 0.5393219685956089, \
    0.5804988662131519, 0.5438674309672835, \
    0.5778781038374717, 0.5483923255655733, \
    0.5752808988764045, 0.5528968376866776, \
    0.5727069351230425, 0.5573811501340064, \
    0.5701559020044543, 0.5618454432626918, \
    0.5676274944567627, 0.5662898950231159, \
    0.565121412803532, 0.5707146810034716, \
    0.5626373626373626, 0.575119974471388, \
    0.5601750547045952, 0.5795059464146423, \
    0.5577342047930284, 0.5838727655809826, \
    0.5553145336225597, 0.588220598517086, \
    0.5529157667386609, 0.5925496096066716, \
    0.5505376344086022, 0.5968599611077938, \
    0.5481798715203426, 0.6011518131893347, \
    0.5458422174840085, 0.6054253239667169, \
    0.5435244161358811, 0.6096806495368553, \
    0.5412262156448203, 0.6139179440123704, \
    0.5389473684210526, 0.6181373595550788, \
    0.5366876310272537, 0.6223390464087787, \
    0.534446764091858, 0.6265231529313529, \
    0.5322245322245323, 0.6306898256261987, \
    0.5300207039337475, 0.6348392091730102, \
    0.5278350515463918, 0.6389714464579207, \
    0.5256673511293635, 0.6430866786030273, \
    0.523517382413088, 0.6471850449953095, \
    0.5213849287169042, 0.6512666833149582, \
    0.5192697768762677, 0.6553317295631277, \
    0.5171717171717172, 0.6593803180891278, \
    0.5150905432595574, 0.6634125816170662, \
    0.5130260521042084, 0.6674286512719563, \
    0.5109780439121756, 0.6714286566053024, \
    0.5089463220675944, 0.6754127256201768, \
    0.5069306930693069, 0.6793809847957973, \
    0.504930966469428, 0.6833335591116206, \
    0.5029469548133595, 0.6872705720709603, \
    0.5009784735812133, 0.691192145724142 }
#if defined(__CUDACC__)
__device__ const double g_logtab[256] = CEL_LOG_TAB;
#endif
#if !defined(__CUDACC__)
static const double h_logtab[256] = CEL_LOG_TAB;
#endif
CEL_HD double log_tab(double E, const double* tab) {
#if defined(__CUDA_ARCH__)
    const int hi = __double2hiint(E);
#else
    long long bits;
    memcpy(&bits, &E, 8);
    const int hi = (int)(bits >> 32);
#endif
    if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return log(E);      // E <= 0, subnormal, Inf, NaN
    const int idx = (hi >> 13) & 0x7f;
    const int k = (hi >> 20) - 1023;
#if defined(__CUDA_ARCH__)
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(E));
#else
    long long mb = (bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL;
    double m;
    memcpy(&m, &mb, 8);
#endif
    const double inv = tab[2 * idx], logc = tab[2 * idx + 1];
    const double r = fma(m, inv, -1.0);
    double p = fma(r, 1.0 / 7.0, -1.0 / 6.0);
    p = fma(p, r, 0.2);
    p = fma(p, r, -0.25);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    p = fma(p, r * r, r);
    return fma((double)k, 0.693147180559945309417232121458, logc + p);
}

// ---------------------------------------------------------------------------------------------
// Cubic B-spline (Interpolations.jl BSpline(Cubic(Line())), OnGrid; un-vendored dependency,
// REQUIRE:21): weights at fractional offset f for taps i-1..i+2, with first/second derivatives.
template <int MODE>
CEL_HD void cubic_weights(double f, double* w, double* dw, double* ddw) {
    const double o = 1.0 - f;
    const double f2 = f * f, o2 = o * o;
    w[0] = (1.0 / 6.0) * o2 * o;
    w[1] = 2.0 / 3.0 - f2 + 0.5 * f2 * f;
    w[2] = 2.0 / 3.0 - o2 + 0.5 * o2 * o;
    w[3] = (1.0 / 6.0) * f2 * f;
    if (MODE >= 1) {
        dw[0] = -0.5 * o2;
        dw[1] = f * (1.5 * f - 2.0);
        dw[2] = o * (2.0 - 1.5 * o);
        dw[3] = 0.5 * f2;
    }
    if (MODE >= 2) {
        ddw[0] = o;
        ddw[1] = 3.0 * f - 2.0;
        ddw[2] = 3.0 * o - 2.0;
        ddw[3] = f;
    }
}

// Star density of fsm_util.jl:225-248 in RAW coordinates: value f0, gradient g0 (2) and
// Hessian h0 (xx, xy, yy) with respect to the spline ARGUMENT (which moves by -J dpos, the
// same Jacobian as the galaxy's x: the epilogue applies it).  coefs: padded (n1 x n2) col-major.
template <int MODE, typename LD>
CEL_HD void star_eval(LD ld, const double* coefs, int n1, int n2, double ax, double ay, double& f0, double* g0,
                      double* h0) {
    const int s1 = n1 - 2, s2 = n2 - 2;
    int ix = (int)floor(ax);
    ix = ix < 1 ? 1 : (ix > s1 - 1 ? s1 - 1 : ix);
    int iy = (int)floor(ay);
    iy = iy < 1 ? 1 : (iy > s2 - 1 ? s2 - 1 : iy);
    double wx[4], dwx[4], ddwx[4], wy[4], dwy[4], ddwy[4];
    cubic_weights<MODE>(ax - ix, wx, dwx, ddwx);
    cubic_weights<MODE>(ay - iy, wy, dwy, ddwy);
    double v = 0, gx = 0, gy = 0, hxx = 0, hxy = 0, hyy = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double* col = coefs + (size_t)(iy - 1 + b) * n1 + (ix - 1);
        const double c0 = ld(col), c1 = ld(col + 1), c2 = ld(col + 2), c3 = ld(col + 3);
        const double r = wx[0] * c0 + wx[1] * c1 + wx[2] * c2 + wx[3] * c3;
        v += wy[b] * r;
        if (MODE >= 1) {
            const double rd = dwx[0] * c0 + dwx[1] * c1 + dwx[2] * c2 + dwx[3] * c3;
            gx += wy[b] * rd;
            gy += dwy[b] * r;
            if (MODE >= 2) {
                const double rdd = ddwx[0] * c0 + ddwx[1] * c1 + ddwx[2] * c2 + ddwx[3] * c3;
                hxx += wy[b] * rdd;
                hxy += dwy[b] * rd;
                hyy += ddwy[b] * r;
            }
        }
    }
    // softpluslikeinv, fsm_util.jl:222
    if (v < 0) {
        const double e = 1e-3 * exp_nonpos(v);
        f0 = e;
        if (MODE >= 1) {
            g0[0] = e * gx;
            g0[1] = e * gy;
        }
        if (MODE >= 2) {
            h0[0] = e * (gx * gx + hxx);
            h0[1] = e * (gx * gy + hxy);
            h0[2] = e * (gy * gy + hyy);
        }
    } else {
        f0 = 1e-3 * (v + 1.0);
        if (MODE >= 1) {
            g0[0] = 1e-3 * gx;
            g0[1] = 1e-3 * gy;
        }
        if (MODE >= 2) {
            h0[0] = 1e-3 * hxx;
            h0[1] = 1e-3 * hxy;
            h0[2] = 1e-3 * hyy;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Galaxy mixture (populate_gal_fsm! fsm_util.jl:194-219 + accum_galaxy_pos!:255-346) in raw
// coordinates.  Component record: mu1 mu2 L11 L12 L22 z   (z excludes gal_frac_dev).
//   r[6]   = d f1 / d(x1, x2, S11, S12, S22, theta)
//   R[21]  = d2 f1 / dy dy packed upper triangle (theta-theta entry is identically 0)
struct GalRaw {
    double f;
    double r[6];
    double R[21];
};

// One group (DEV: the 8 de Vaucouleurs prototypes, else the 6 exponential ones) of the mixture.
// KT > 0 fixes the PSF component count at compile time so the k loop unrolls and the K exp chains of
// one prototype interleave; KT == 0 takes K at run time.
struct GalAcc {
    double f, ft, ax1, ax2, tx1, tx2, u1, u2, u3, as1, as2, as3, ts1, ts2, ts3;
    double xs11, xs12, xs13, xs21, xs22, xs23, ss11, ss12, ss13, ss22, ss23, ss33;
};

template <int MODE, int KT, bool DEV, typename LD>
CEL_HD void gal_group(LD ld, const double* comps, int Krt, const double* nu, const double* etab, double thc, double hx,
                      double wy, GalAcc& A) {
    const int K = KT > 0 ? KT : Krt;
    const int j0 = DEV ? 0 : NPROTO_DEV, j1 = DEV ? NPROTO_DEV : NPROTO;
    for (int j = j0; j < j1; ++j) {
        const double nuc = nu[j];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double* cp = comps + (j * K + k) * COMP_STRIDE;
            const double mu1 = ld(cp), mu2 = ld(cp + 1), l11 = ld(cp + 2), l12 = ld(cp + 3), l22 = ld(cp + 4),
                         z = ld(cp + 5);
            const double d1 = hx - mu1, d2 = wy - mu2;
            const double p1 = l11 * d1 + l12 * d2;
            const double p2 = l12 * d1 + l22 * d2;
            const double q = d1 * p1 + d2 * p2;
            const double fp = z * exp_scaled_tab(q, -0.5, etab);   // f_pre, BivariateNormals.jl:219
            if (MODE == 1) {
                // gradient only: accumulate UNWEIGHTED first-order sums; the caller scales the group by
                // theta_i once and obtains d/dtheta from the difference of the two groups (fsm_util.jl:277-291)
                const double hl11 = ld(cp + 6), hl22 = ld(cp + 7);
                A.f += fp;
                A.ax1 += fp * p1;
                A.ax2 += fp * p2;
                const double fn = fp * nuc;
                A.as1 += fn * fma(0.5 * p1, p1, -hl11);     // bvn_sig_d, BivariateNormals.jl:267-272
                A.as2 += fn * fma(p1, p2, -l12);
                A.as3 += fn * fma(0.5 * p2, p2, -hl22);
                continue;
            }
            const double w = thc * fp;
            A.f += w;
            if (MODE >= 1) {
                const double wd = DEV ? fp : -fp;    // gal_frac_dev_dir * f_pre, fsm_util.jl:291
                A.ft += wd;
                A.ax1 += w * p1;
                A.ax2 += w * p2;
                const double hl11 = 0.5 * l11, hl22 = 0.5 * l22;   // (records 6, 7 hold them; cheaper to recompute here)
                const double a = p1 * p1, b = p1 * p2, cc = p2 * p2;
                const double g1 = fma(0.5, a, -hl11);    // bvn_sig_d, BivariateNormals.jl:267-272
                const double g2 = b - l12;
                const double g3 = fma(0.5, cc, -hl22);
                const double wn = w * nuc;
                A.as1 += wn * g1;
                A.as2 += wn * g2;
                A.as3 += wn * g3;
                if (MODE >= 2) {
                    const double wdn = wd * nuc;
                    A.tx1 += wd * p1;
                    A.tx2 += wd * p2;
                    A.ts1 += wdn * g1;
                    A.ts2 += wdn * g2;
                    A.ts3 += wdn * g3;
                    A.u1 += w * g1;
                    A.u2 += w * g2;
                    A.u3 += w * g3;
                    // bvn_xsig_h (BivariateNormals.jl:310-316) + g_x g_S', with g_x = -p
                    A.xs11 += wn * (p1 * (l11 - g1));
                    A.xs12 += wn * (p1 * (l12 - g2) + p2 * l11);
                    A.xs13 += wn * (p2 * l12 - p1 * g3);
                    A.xs21 += wn * (p1 * l12 - p2 * g1);
                    A.xs22 += wn * (p2 * (l12 - g2) + p1 * l22);
                    A.xs23 += wn * (p2 * (l22 - g3));
                    // bvn_sigsig_h (BivariateNormals.jl:293-306, dsiginv_dsig:168-183) + g_S g_S'
                    const double wnn = wn * nuc;
                    A.ss11 += wnn * (l11 * (hl11 - a) + g1 * g1);
                    A.ss12 += wnn * (l12 * (l11 - a) - b * l11 + g1 * g2);
                    A.ss13 += wnn * (l12 * (0.5 * l12 - b) + g1 * g3);
                    A.ss22 += wnn * (l22 * (l11 - a) + l12 * (l12 - 2.0 * b) - cc * l11 + g2 * g2);
                    A.ss23 += wnn * (l12 * (l22 - cc) - b * l22 + g2 * g3);
                    A.ss33 += wnn * (l22 * (hl22 - cc) + g3 * g3);
                }
            }
        }
    }
}

template <int MODE, int KT, typename LD>
CEL_HD void gal_eval(LD ld, const double* comps, int K, const double* nu /*14*/, const double* etab, double theta,
                     double hx, double wy, GalRaw& o) {
    GalAcc A;
    A.f = A.ft = A.ax1 = A.ax2 = A.tx1 = A.tx2 = A.u1 = A.u2 = A.u3 = 0.0;
    A.as1 = A.as2 = A.as3 = A.ts1 = A.ts2 = A.ts3 = 0.0;
    A.xs11 = A.xs12 = A.xs13 = A.xs21 = A.xs22 = A.xs23 = 0.0;
    A.ss11 = A.ss12 = A.ss13 = A.ss22 = A.ss23 = A.ss33 = 0.0;
    if (MODE == 1) {
        gal_group<MODE, KT, true>(ld, comps, K, nu, etab, theta, hx, wy, A);
        GalAcc Bx;
        Bx.f = Bx.ax1 = Bx.ax2 = Bx.as1 = Bx.as2 = Bx.as3 = 0.0;
        gal_group<MODE, KT, false>(ld, comps, K, nu, etab, 1.0 - theta, hx, wy, Bx);
        const double t0 = theta, t1 = 1.0 - theta;
        o.f = t0 * A.f + t1 * Bx.f;
        o.r[0] = -(t0 * A.ax1 + t1 * Bx.ax1);
        o.r[1] = -(t0 * A.ax2 + t1 * Bx.ax2);
        o.r[2] = t0 * A.as1 + t1 * Bx.as1;
        o.r[3] = t0 * A.as2 + t1 * Bx.as2;
        o.r[4] = t0 * A.as3 + t1 * Bx.as3;
        o.r[5] = A.f - Bx.f;
        return;
    }
    gal_group<MODE, KT, true>(ld, comps, K, nu, etab, theta, hx, wy, A);
    gal_group<MODE, KT, false>(ld, comps, K, nu, etab, 1.0 - theta, hx, wy, A);
    const double f = A.f, ft = A.ft, ax1 = A.ax1, ax2 = A.ax2, tx1 = A.tx1, tx2 = A.tx2, u1 = A.u1, u2 = A.u2,
                 u3 = A.u3, as1 = A.as1, as2 = A.as2, as3 = A.as3, ts1 = A.ts1, ts2 = A.ts2, ts3 = A.ts3;
    const double xs11 = A.xs11, xs12 = A.xs12, xs13 = A.xs13, xs21 = A.xs21, xs22 = A.xs22, xs23 = A.xs23;
    const double ss11 = A.ss11, ss12 = A.ss12, ss13 = A.ss13, ss22 = A.ss22, ss23 = A.ss23, ss33 = A.ss33;
    o.f = f;
    if (MODE >= 1) {
        o.r[0] = -ax1;
        o.r[1] = -ax2;
        o.r[2] = as1;
        o.r[3] = as2;
        o.r[4] = as3;
        o.r[5] = ft;
    }
    if (MODE >= 2) {
        double* R = o.R;
        // H_xx + g_x g_x' = (p1^2 - L11, p1 p2 - L12, p2^2 - L22) = (2 g1, g2, 2 g3)
        R[tri6(0, 0)] = 2.0 * u1;
        R[tri6(0, 1)] = u2;
        R[tri6(1, 1)] = 2.0 * u3;
        R[tri6(0, 2)] = xs11;
        R[tri6(0, 3)] = xs12;
        R[tri6(0, 4)] = xs13;
        R[tri6(1, 2)] = xs21;
        R[tri6(1, 3)] = xs22;
        R[tri6(1, 4)] = xs23;
        R[tri6(0, 5)] = -tx1;
        R[tri6(1, 5)] = -tx2;
        R[tri6(2, 2)] = ss11;
        R[tri6(2, 3)] = ss12;
        R[tri6(2, 4)] = ss13;
        R[tri6(3, 3)] = ss22;
        R[tri6(3, 4)] = ss23;
        R[tri6(4, 4)] = ss33;
        R[tri6(2, 5)] = ts1;
        R[tri6(3, 5)] = ts2;
        R[tri6(4, 5)] = ts3;
        R[tri6(5, 5)] = 0.0;
    }
}

// value-only mixture for a neighbour (is_active_source == false, fsm_util.jl:265)
template <int KT, typename LD>
CEL_HD double gal_value(LD ld, const double* comps, int Krt, const double* etab, double theta, double hx, double wy) {
    const int K = KT > 0 ? KT : Krt;
    double fg[2] = {0.0, 0.0};
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const int j0 = g == 0 ? 0 : NPROTO_DEV, j1 = g == 0 ? NPROTO_DEV : NPROTO;
        for (int j = j0; j < j1; ++j) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double* cp = comps + (j * K + k) * COMP_STRIDE;
                const double mu1 = ld(cp), mu2 = ld(cp + 1), l11 = ld(cp + 2), l12 = ld(cp + 3), l22 = ld(cp + 4),
                             z = ld(cp + 5);
                const double d1 = hx - mu1, d2 = wy - mu2;
                const double p1 = l11 * d1 + l12 * d2;
                const double p2 = l12 * d1 + l22 * d2;
                fg[g] += z * exp_scaled_tab(d1 * p1 + d2 * p2, -0.5, etab);
            }
        }
    }
    return theta * fg[0] + (1.0 - theta) * fg[1];
}

// ---------------------------------------------------------------------------------------------
// Pixel term (add_elbo_log_term! elbo_objective.jl:274-327 + add_scaled_sfs! :383-385 + :391)
// as a function of z = (A1, A2, B1, B2, f0, f1), and its accumulation in (c, y) space.
//   E  = Ebg + A1 f0 + A2 f1
//   V  = Vbg + B1 f0^2 + B2 f1^2 - (A1 f0 + A2 f1)^2         (calculate_G_s! :204)
//   L  = x (log E - V / (2 E^2)) - iota E + pixconst,  pixconst = x log(iota) - lgamma(x + 1)
struct PixelConsts {
    double x, iota, pixconst;
};

// acc: per-thread accumulator slots, element a at acc[a * stride]
template <int MODE>
CEL_HD void pixel_accumulate(double* acc, int stride, const PixelConsts& pc, double Ebg, double Vbg, bool covered,
                             bool add_value, const double* cb /*A1 A2 B1 B2*/, double f0, const double* g0,
                             const double* h0, const GalRaw& gal, const double* logtab = nullptr) {
    const double A1 = cb[0], A2 = cb[1], B1 = cb[2], B2 = cb[3];
    const double f1 = gal.f;
    const double m = covered ? (A1 * f0 + A2 * f1) : 0.0;
    const double E = Ebg + m;
    const double V = covered ? (Vbg + B1 * f0 * f0 + B2 * f1 * f1 - m * m) : Vbg;
    const double iE = 1.0 / E;
    const double iE2 = iE * iE;
    if (add_value) acc[ACC_VAL * stride] += pc.x * ((logtab ? log_tab(E, logtab) : log(E)) - 0.5 * V * iE2) - pc.iota * E + pc.pixconst;
    if (MODE == 0 || !covered) return;

    const double gE = pc.x * (iE + V * iE2 * iE) - pc.iota;     // combine_grad[2] * x - iota
    const double gV = -0.5 * pc.x * iE2;                         // combine_grad[1] * x
    const double Ez[6] = {f0, f1, 0.0, 0.0, A1, A2};
    const double Vz[6] = {-2.0 * m * f0, -2.0 * m * f1, f0 * f0, f1 * f1, 2.0 * (B1 * f0 - m * A1),
                          2.0 * (B2 * f1 - m * A2)};
    double Lz[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) Lz[i] = gE * Ez[i] + gV * Vz[i];

    // first order: G[k] = L_f1 r[k] + L_f0 g0[k];   C1[c] = L_c
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double g = Lz[5] * gal.r[k];
        if (k < 2) g += Lz[4] * g0[k];
        acc[(ACC_G + k) * stride] += g;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[(ACC_C1 + c) * stride] += Lz[c];
    if (MODE == 1) return;

    const double LEE = -pc.x * (iE2 + 3.0 * V * iE2 * iE2);      // combine_hess[2,2] * x
    const double LEV = pc.x * iE2 * iE;                           // combine_hess[1,2] * x
    double Lzz[6][6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 6; ++j) {
            double ezz = 0.0, bpart = 0.0;
            if ((i == 0 && j == 4) || (i == 1 && j == 5)) ezz = 1.0;
            if (i == 2 && j == 4) bpart = 2.0 * f0;
            if (i == 4 && j == 4) bpart = 2.0 * B1;
            if (i == 3 && j == 5) bpart = 2.0 * f1;
            if (i == 5 && j == 5) bpart = 2.0 * B2;
            const double vzz = -2.0 * (Ez[i] * Ez[j] + m * ezz) + bpart;
            Lzz[i][j] = LEE * Ez[i] * Ez[j] + LEV * (Ez[i] * Vz[j] + Vz[i] * Ez[j]) + gE * ezz + gV * vzz;
        }
    // HH[k][l] = L_f1 R[k][l] + L_f1f1 r_k r_l + (x block) L_f0 h0 + L_f0f0 g0 g0' + L_f0f1 (g0 r' + r g0')
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int l = k; l < 6; ++l) {
            double v = Lz[5] * gal.R[tri6(k, l)] + Lzz[5][5] * gal.r[k] * gal.r[l];
            if (k < 2) v += Lzz[4][5] * g0[k] * gal.r[l];
            if (l < 2) {
                v += Lzz[4][5] * gal.r[k] * g0[l];
                v += Lz[4] * h0[k + l] + Lzz[4][4] * g0[k] * g0[l];   // h0 packed xx, xy, yy
            }
            acc[(ACC_HH + tri6(k, l)) * stride] += v;
        }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c; d < 4; ++d) acc[(ACC_CC + tri4(c, d)) * stride] += Lzz[c][d];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double v = Lzz[c][5] * gal.r[k];
            if (k < 2) v += Lzz[c][4] * g0[k];
            acc[(ACC_CR + c * 6 + k) * stride] += v;
        }
}

// ---------------------------------------------------------------------------------------------
// Per-(source, image) mixture set-up: load_bvn_mixtures! (fsm_util.jl:111-169), GalaxyCacheComponent
// (:37-65), BvnComponent (BivariateNormals.jl:151-191), get_bvn_cov (:29-43); one call per component.
// psf7: alphaBar, xiBar[2], tauBar col-major (4).  Writes a 6-double component record.
// XiXi(rho, phi, sigma) of get_bvn_cov (BivariateNormals.jl:29-43): the same for every component and image of a source
CEL_HD void galaxy_xixi(double rho, double phi, double sigma, double& x11, double& off, double& x22) {
    const double cp = cos(phi), sp = sin(phi);
    const double ab_term = rho * rho - 1.0;
    const double ss = sigma * sigma;
    off = -ss * cp * sp * ab_term;
    x11 = ss * (1.0 + ab_term * (sp * sp));
    x22 = ss * (1.0 + ab_term * (cp * cp));
}
CEL_HD void make_component_xi(const double* psf7, double eta, double nuBar, double m1, double m2, double x11, double off,
                              double x22, double* out) {
    const double v11 = psf7[3] + nuBar * x11;
    const double v21 = psf7[4] + nuBar * off;
    const double v12 = psf7[5] + nuBar * off;
    const double v22 = psf7[6] + nuBar * x22;
    const double det = v11 * v22 - v12 * v21;
    const double idet = 1.0 / det;
    out[0] = psf7[1] + m1;
    out[1] = psf7[2] + m2;
    out[2] = v22 * idet;
    out[3] = -v12 * idet;
    out[4] = v11 * idet;
    out[5] = (psf7[0] * eta) * (1.0 / (sqrt(det) * 6.283185307179586476925286766559));
    out[6] = 0.5 * out[2];
    out[7] = 0.5 * out[4];
}
CEL_HD void make_component(const double* psf7, double eta, double nuBar, double m1, double m2, double rho, double phi,
                           double sigma, double* out) {
    double x11, off, x22;
    galaxy_xixi(rho, phi, sigma, x11, off, x22);
    make_component_xi(psf7, eta, nuBar, m1, m2, x11, off, x22, out);
}

// GalaxySigmaDerivs (BivariateNormals.jl:346-397) with nuBar = 1: J0[k][j] = dSigma_k/dshape_j,
// T0[k][j][l] = d2 Sigma_k / dshape_j dshape_l, shape = (axis_ratio, angle, radius).
CEL_HD void sigma_derivs(double rho, double phi, double sigma, double J0[3][3], double T0[3][3][3]) {
    const double c = cos(phi), s = sin(phi);
    const double cs = c * s, s2 = s * s, c2 = c * c;
    const double rr = sigma * sigma;
    const double ab_term = rho * rho - 1.0;
    const double X11 = rr * (1.0 + ab_term * s2), X12 = -rr * cs * ab_term, X22 = rr * (1.0 + ab_term * c2);
    const double a1 = 2.0 * rho * rr;
    J0[0][0] = a1 * s2;
    J0[1][0] = -a1 * cs;
    J0[2][0] = a1 * c2;
    const double a2 = rr * ab_term;
    J0[0][1] = a2 * (2.0 * cs);
    J0[1][1] = a2 * (s2 - c2);
    J0[2][1] = a2 * (-2.0 * cs);
    J0[0][2] = 2.0 * X11 / sigma;
    J0[1][2] = 2.0 * X12 / sigma;
    J0[2][2] = 2.0 * X22 / sigma;
    const double t2 = 2.0 * rr;
    T0[0][0][0] = s2 * t2;
    T0[1][0][0] = -cs * t2;
    T0[2][0][0] = c2 * t2;
    const double t3 = t2 * rho;
    T0[0][1][0] = T0[0][0][1] = 2.0 * cs * t3;
    T0[1][1][0] = T0[1][0][1] = (s2 - c2) * t3;
    T0[2][1][0] = T0[2][0][1] = -2.0 * cs * t3;
    const double t4 = t2 * ab_term;
    T0[0][1][1] = (c2 - s2) * t4;
    T0[1][1][1] = 2.0 * cs * t4;
    T0[2][1][1] = (s2 - c2) * t4;
    for (int k = 0; k < 3; ++k) {
        T0[k][2][0] = T0[k][0][2] = 2.0 * J0[k][0] / sigma;
        T0[k][2][1] = T0[k][1][2] = 2.0 * J0[k][1] / sigma;
    }
    T0[0][2][2] = 2.0 * X11 / rr;
    T0[1][2][2] = 2.0 * X12 / rr;
    T0[2][2][2] = 2.0 * X22 / rr;
}

// Source brightness (source_brightness.jl:27-202): E_l[b][i] = exp(kappa_b . beta_i), E_ll[b][i] =
// exp(lambda_b . beta_i) over beta = (flux_loc, flux_scale, color_mean 1..4, color_var 1..4).
// Every derivative the reference builds with multiply_sfs! is E * kappa (x) kappa.
CEL_HD void band_coefs(int b /*0..4*/, double kappa[10], double lambda[10]) {
    for (int k = 0; k < 10; ++k) kappa[k] = lambda[k] = 0.0;
    kappa[0] = 1.0;
    kappa[1] = 0.5;
    lambda[0] = 2.0;
    lambda[1] = 2.0;
    // colours: band 4 uses c3; band 5 c3,c4; band 2 uses -c2; band 1 uses -c2,-c1
    const int use[5][4] = {{-1, -1, 0, 0}, {0, -1, 0, 0}, {0, 0, 0, 0}, {0, 0, 1, 0}, {0, 0, 1, 1}};
    for (int m = 0; m < 4; ++m) {
        const int u = use[b][m];
        if (u != 0) {
            kappa[2 + m] = (double)u;
            kappa[6 + m] = 0.5;
            lambda[2 + m] = 2.0 * u;
            lambda[6 + m] = 2.0;
        }
    }
}
// canonical (0-based) index of brightness parameter k (0..9) of type i: brightness_standard_alignment
CEL_HD int bright_id(int i, int k) {
    return k == 0 ? 6 + i : (k == 1 ? 8 + i : (k < 6 ? 10 + (k - 2) + 4 * i : 18 + (k - 6) + 4 * i));
}
CEL_HD void brightness_values(const double* vs, double El[2][5], double Ell[2][5]) {
    for (int i = 0; i < 2; ++i) {
        double beta[10];
        for (int k = 0; k < 10; ++k) beta[k] = vs[bright_id(i, k)];
        for (int b = 0; b < 5; ++b) {
            double ka[10], la[10];
            band_coefs(b, ka, la);
            double s1 = 0, s2 = 0;
            for (int k = 0; k < 10; ++k) {
                s1 += ka[k] * beta[k];
                s2 += la[k] * beta[k];
            }
            El[i][b] = exp(s1);
            Ell[i][b] = exp(s2);
        }
    }
}

}  // namespace celeste
#endif
