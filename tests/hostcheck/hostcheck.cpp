// TEST-ONLY: compiles the host/device headers of tft_vs_fund_b200/csrc with g++
// so the thread-level math that the CUDA kernels inline can be checked against
// the oracle on a machine without a GPU.  Not linked into the product library,
// never on the product path.
#include "tvf_pose.cuh"
#include "tvf_scene.cuh"

using namespace tvf;

extern "C" {

void hc_null3(const double* M, double* v) { null3(M, v); }
// the QR + inverse-iteration route alone: 1 = it answered (v valid), 0 = it declined (null3 then takes the Jacobi route)
int hc_null3_qr(const double* M, int transpose, double* v) { return null3_qr(M, transpose != 0, v) ? 1 : 0; }

int hc_jacobi_sweeps(const double* M) {
    double A[9], V[9], s[3];
    for (int i = 0; i < 9; ++i) A[i] = M[i];
    int sweeps = 0;
    jacobi_svd3(A, V, s, &sweeps);
    return sweeps;
}

void hc_svd3(const double* M, double* U, double* s, double* V) { svd3_full(M, U, s, V); }

void hc_inv3(const double* M, double* Mi) { inv3(M, Mi); }

void hc_transform_tft(const double* T, const double* M1, const double* M2, const double* M3, int inverse, double* Tn) {
    transform_tft(T, M1, M2, M3, inverse, Tn);
}

void hc_tft_from_p(const double* P1, const double* P2, const double* P3, double* T) { tft_from_p(P1, P2, P3, T); }

void hc_epipoles(const double* T, double* e21, double* e31) { tft_epipoles(T, e21, e31); }

void hc_onb3(const double* e, double* u1, double* u2) { onb3(e, u1, u2); }

void hc_ang_error(const double* a, const double* b, double* r, double* t) { ang_error(a, b, r, t); }

// rows: M x 4 row-major
int hc_dlt(const double* rows, int M, double* x) {
    int it = 0;
    if (M == 4) {
        double a[4][4];
        for (int i = 0; i < 16; ++i) a[i / 4][i % 4] = rows[i];
        dlt_null<4>(a, x, &it);
    } else {
        double a[6][4];
        for (int i = 0; i < 24; ++i) a[i / 4][i % 4] = rows[i];
        dlt_null<6>(a, x, &it);
    }
    return it;
}

// Certified depth signs (route 0: dlt4_depth_signs, route 1: dlt4_depth_signs_ray) against the accurate route on every point and candidate of a problem:
// counts[0] = DLTs examined, counts[1] = answered by the shortcut, counts[2] = shortcut answers that DIFFER from the
// accurate route's signs (must stay 0).
void hc_vote_signs_check(int route, int mode, const double* model, const double* CalM, const double* corresp, int n, long long* counts) {
    double cand[CAND_SIZE];
    if (mode == 0) candidates_from_tft(model, CalM, cand); else candidates_from_f(model, model + 9, CalM, cand);
    double P1[12];
    load_K1_as_P1(CalM, P1);
    for (int i = 0; i < n; ++i) {
        const double* p = corresp + 6 * i;
        double ra[4], rb[4];
        dlt_rows(P1, p[0], p[1], ra, rb);
        for (int pair = 0; pair < 2; ++pair)
            for (int q = 0; q < 2; ++q) {
                double P[12], r3[3], tz;
                candidate_camera(cand + pair * CAND_PAIR, q == 0 ? 0 : 3, P, r3, &tz);
                double a[4][4], b[4][4];
                for (int e = 0; e < 4; ++e) { a[0][e] = ra[e]; a[1][e] = rb[e]; }
                dlt_rows(P, p[2 + 2 * pair], p[3 + 2 * pair], a[2], a[3]);
                for (int r = 0; r < 4; ++r) for (int e = 0; e < 4; ++e) b[r][e] = a[r][e];
                counts[0] += 1;
                int sx, sz;
                if (route == 1) {                  // certified ray / plane test (dlt4_depth_signs_ray), general form
                    double m7[7];
                    dlt_row_minors(ra, rb, m7);
                    if (!dlt4_depth_signs_ray(m7, a[2], a[3], r3, tz, &sx, &sz)) continue;
                } else if (route == 2) {           // the form for a view-1 camera K1*[I | 0] (what the kernels run)
                    double m7[7];
                    dlt_row_minors<true>(ra, rb, m7);
                    if (!dlt4_depth_signs_ray<true>(m7, a[2], a[3], r3, tz, &sx, &sz)) continue;
                } else if (!dlt4_depth_signs(a, r3, tz, &sx, &sz)) continue;
                counts[1] += 1;
                double X[4];
                dlt_null<4>(b, X);
                const double iw = 1.0 / X[3];
                const double X0 = X[0] * iw, X1 = X[1] * iw, X2 = X[2] * iw, X3 = X[3] * iw;
                const double z2 = r3[0] * X0 + r3[1] * X1 + r3[2] * X2 + tz * X3;
                if ((int)sign_(X2) != sx || (int)sign_(z2) != sz) counts[2] += 1;
            }
    }
}

// Whole pose tail on one problem, stage by stage exactly as the kernels chain them.
// mode 0: `model` is T (27, pixel coordinates); mode 1: `model` is [F21(9) F31(9)].
int hc_pose_tail(int mode, const double* model, const double* CalM, const double* corresp, int n,
                 double* Rt2, double* Rt3, double* reconst, double* repr, int* votes8) {
    double cand[CAND_SIZE];
    int st = (mode == 0) ? candidates_from_tft(model, CalM, cand) : candidates_from_f(model, model + 9, CalM, cand);
    double P1[12];
    load_K1_as_P1(CalM, P1);
    int vote[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int nan2 = 0, nan3 = 0;
    {
        int v2[2] = {0, 0}, v3[2] = {0, 0}, n2 = 0, n3 = 0;
        for (int i = 0; i < n; ++i) {
            const double* p = corresp + 6 * i;
            double ra[4], rb[4];
            dlt_rows(P1, p[0], p[1], ra, rb);
            double m7[7];
            dlt_row_minors<true>(ra, rb, m7);
            cheirality_point<true>(ra, rb, m7, cand, p[2], p[3], v2, &n2, nullptr, nullptr);
            cheirality_point<true>(ra, rb, m7, cand + CAND_PAIR, p[4], p[5], v3, &n3, nullptr, nullptr);
        }
        expand_votes(v2, n2, vote, &nan2);
        expand_votes(v3, n3, vote + 4, &nan3);
    }
    for (int k = 0; k < 8; ++k) votes8[k] = vote[k];
    const int k2 = select_candidate(vote, nan2), k3 = select_candidate(vote + 4, nan3);
    if (k2 < 0) st |= ST_NO_POSE_2;
    if (k3 < 0) st |= ST_NO_POSE_3;
    if (k2 < 0 || k3 < 0) return st;
    double P2[12], P3[12];
    selected_pose(cand, k2, Rt2, P2);
    selected_pose(cand + CAND_PAIR, k3, Rt3, P3);
    double num = 0.0, den = 0.0;
    for (int i = 0; i < n; ++i) {
        double a, b;
        scale_point(P1, P2, P3, P3 + 9, corresp + 6 * i, &a, &b);
        num += a; den += b;
    }
    const double lam = -num / den;
    for (int i = 0; i < 3; ++i) { Rt3[9 + i] *= lam; P3[9 + i] *= lam; }
    double sq = 0.0;
    for (int i = 0; i < n; ++i) sq += final_point(P1, P2, P3, corresp + 6 * i, reconst + 3 * i);
    *repr = sqrt(sq / (3.0 * n));
    return st;
}

// one sweep trial on the host with the very code the device kernel runs (RNG streams exposed for checks)
void hc_scene_trial(const double* P, int n, double noise, unsigned seed, double hi_x, double hi_y, double* out) {
    static MT19937 rng;
    static double c[6 * SCENE_MAX_POINTS];
    static unsigned char arr[SCENE_MAX_POINTS];
    static signed char outpos[SCENE_MAX_POINTS];
    scene_trial(rng, P, n, noise, seed, hi_x, hi_y, out, c, arr, outpos);
}

void hc_scene_seed(const double* P, int n, const double* noise_levels, int lv_lo, int lv_hi, unsigned seed, double hi_x,
                   double hi_y, double* out) {
    static MT19937 rng, snap;
    static double clean[6 * SCENE_MAX_POINTS], z[6 * SCENE_MAX_POINTS], c[6 * SCENE_MAX_POINTS];
    unsigned char arr[SCENE_MAX_POINTS];
    signed char outpos[SCENE_MAX_POINTS];
    scene_seed_levels(rng, snap, P, n, noise_levels, lv_lo, lv_hi, seed, hi_x, hi_y, out, clean, z, c, arr, outpos);
}

void hc_mt_streams(unsigned seed, int n, double* uniform, double* normal, unsigned* ints, unsigned maxv) {
    MT19937 r;
    r.seed(seed);
    for (int i = 0; i < n; ++i) uniform[i] = r.res53();
    for (int i = 0; i < n; ++i) normal[i] = r.normal();
    for (int i = 0; i < n; ++i) ints[i] = r.interval(maxv);
}

}  // extern "C"
