// Host build of the exact minimal solver (zebrapose_b200/csrc/zp_cvepnp.cuh): stage A and C run as plain serial code
// (stride-1 views), the six lanes of stage B are emulated by a loop with the barriers at the loop boundaries.  Lets the
// CPU suite check the product source bit for bit against the oracle (oracle/cv_epnp.c) and cv2.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include <vector>
#include "../../zebrapose_b200/csrc/zp_cvepnp.cuh"

extern "C" int cve_host_epnp(const float* corr, int cap, const int32_t* idx, int m, const double* K4, double* pose12,
                             double* cand39, int* steps) {
    std::vector<double> rec(CVH_DOUBLES, 0.0), wk(96, 0.0), A(12 * CVE_RS, 0.0), W(12, 0.0);
    double* r = rec.data();
    const CveCam cam{K4[0], K4[1], K4[2], K4[3]};
    cve_stage_a(corr, cap, idx, m, cam, cve_dv(r + CVH_PW, 1), cve_dv(r + CVH_US, 1), cve_dv(r + CVH_AL, 1),
                cve_dv(r + CVH_CW, 1), cve_dv(wk.data(), 1));
    // stage B, six emulated lanes
    int flags[4] = {0, 0, 0, 0};
    CveJ j[CVE_G];
    double sums[CVE_G][13];
    for (int l = 0; l < CVE_G; l++) cve_b_mtm_table(A.data(), l, r + CVH_AL, r + CVH_US, m, cam);
    for (int l = 0; l < CVE_G; l++) cve_b_mtm_sums(A.data(), l, r + CVH_AL, r + CVH_US, m, cam, sums[l]);
    for (int l = 0; l < CVE_G; l++) cve_b_mtm_store(A.data(), l, sums[l]);
    for (int l = 0; l < CVE_G; l++) j[l] = cve_j_make(A.data(), CVE_RS, 12, 12, l, CVE_G, flags, true);
    for (int l = 0; l < CVE_G; l++) cve_j_init(j[l]);
    int T = 1;
    for (;; T++) {
        for (int l = 0; l < CVE_G; l++) if (!j[l].done) cve_jstep_a<12, 12>(j[l], T);
        for (int l = 0; l < CVE_G; l++) if (!j[l].done) cve_jstep_c<12>(j[l], T);
        bool all = true;
        for (int l = 0; l < CVE_G; l++) all = all && j[l].done;
        if (all) break;
    }
    if (steps) steps[0] = T;
    for (int l = 0; l < CVE_G; l++) cve_b_norms(A.data(), l, W.data());
    bool fast = true;
    for (int l = 0; l < CVE_G; l++) fast = cve_b_tail(A.data(), l, W.data(), r + CVH_V4) && fast;
    if (!fast) cve_b_finish(A.data(), W.data(), r + CVH_V4);
    for (int l = 0; l < CVE_G; l++) cve_b_L_rho(r + CVH_V4, l, cve_dv(r + CVH_CW, 1), cve_dv(r + CVH_L, 1), cve_dv(r + CVH_RHO, 1));
    for (int c = 0; c < 3; c++)
        cve_stage_c(c, m, cam, cve_dv(r + CVH_L, 1), cve_dv(r + CVH_RHO, 1), cve_dv(r + CVH_V4, 1), cve_dv(r + CVH_AL, 1),
                    cve_dv(r + CVH_PW, 1), cve_dv(r + CVH_US, 1), cve_dv(wk.data(), 1), cve_dv(r + CVH_OUT + 13 * c, 1));
    const int N = cve_pick(cve_dv(r + CVH_OUT, 1));
    for (int i = 0; i < 12; i++) pose12[i] = r[CVH_OUT + 13 * N + i];
    if (cand39) for (int i = 0; i < 39; i++) cand39[i] = r[CVH_OUT + i];
    return N + 1;
}
