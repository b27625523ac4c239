// Host build of the exact minimal solver (zebrapose_b200/csrc/zp_cvepnp.cuh): the six lanes of a hypothesis are emulated
// by a loop, barriers become the boundaries between the loops.  Lets the CPU suite check the product source bit for bit
// against the oracle (oracle/cv_epnp.c) and cv2.  Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include <vector>
#include "../../zebrapose_b200/csrc/zp_cvepnp.cuh"

template <int M, bool HASV>
static int jrun(CveJ* j) {
    for (int l = 0; l < CVE_G; l++) cve_j_init(j[l]);
    int T = 1;
    for (;; T++) {
        for (int l = 0; l < CVE_G; l++) if (!j[l].done) cve_jstep_a<M, HASV>(j[l], T);
        for (int l = 0; l < CVE_G; l++) if (!j[l].done) cve_jstep_c(j[l], T);
        bool all = true;
        for (int l = 0; l < CVE_G; l++) all = all && j[l].done;
        if (all) break;
    }
    return T;
}

extern "C" int cve_host_epnp(const float* corr, int cap, const int32_t* idx, int m, const double* K4, double* pose12,
                             double* cand39, int* steps5) {
    std::vector<double> Sv(CVE_HB, 0.0);
    double* S = Sv.data();
    int* flags = (int*)(S + CVE_FLAGS);
    const CveCam cam{K4[0], K4[1], K4[2], K4[3]};
    CveJ j[CVE_G];
    for (int l = 0; l < CVE_G; l++) cve_ph0(S, l, corr, cap, idx, m, cam);
    for (int l = 0; l < CVE_G; l++) j[l] = l == 0 ? cve_j_make(S + CVE_A3, 3, S + CVE_V3, 3, 3, 3, 0, 1, flags) : cve_j_none();
    steps5[0] = jrun<3, true>(j);
    for (int l = 0; l < CVE_G; l++) cve_ph1(S, l, m);
    for (int l = 0; l < CVE_G; l++) j[l] = l == 0 ? cve_j_make(S + CVE_A3, 3, S + CVE_V3, 3, 3, 3, 0, 1, flags) : cve_j_none();
    steps5[1] = jrun<3, true>(j);
    for (int l = 0; l < CVE_G; l++) cve_ph2(S, l);
    for (int l = 0; l < CVE_G; l++) cve_ph3(S, l, m);
    for (int l = 0; l < CVE_G; l++) cve_ph4(S, l, m, cam);
    for (int l = 0; l < CVE_G; l++) j[l] = cve_j_make(S + CVE_A, CVE_RS, nullptr, 0, 12, 12, l, 6, flags);
    steps5[2] = jrun<12, false>(j);
    for (int l = 0; l < CVE_G; l++) cve_ph5(S, l);
    for (int l = 0; l < CVE_G; l++) cve_ph6(S, l);
    for (int l = 0; l < CVE_G; l++) cve_ph7(S, l);
    for (int l = 0; l < CVE_G; l++) {
        const int c = cve_lane_cand(l);
        if (c < 0) { j[l] = cve_j_none(); continue; }
        const CveCand k = cve_cand(c);
        j[l] = cve_j_make(S + CVE_A + k.at, 6, S + CVE_A + k.vt, k.nc, k.nc, 6, l - k.lane0, k.nl, flags + 4 * c);
    }
    steps5[3] = jrun<6, true>(j);
    double betas[CVE_G][4];
    for (int l = 0; l < CVE_G; l++) cve_ph8(S, l, m, betas[l]);
    for (int l = 0; l < CVE_G; l++) cve_ph9(S, l, m, betas[l]);
    for (int l = 0; l < CVE_G; l++) {
        const int c = cve_lane_cand(l);
        if (c < 0 || l != cve_cand(c).lane0) { j[l] = cve_j_none(); continue; }
        double* slot = S + CVE_A + 48 * c;
        j[l] = cve_j_make(slot + 24, 3, slot + 33, 3, 3, 3, 0, 1, flags + 4 * c);
    }
    steps5[4] = jrun<3, true>(j);
    for (int l = 0; l < CVE_G; l++) cve_ph10(S, l, m, cam);
    const double* o = cve_pick(S);
    for (int i = 0; i < 12; i++) pose12[i] = o[i];
    if (cand39) for (int i = 0; i < 39; i++) cand39[i] = S[CVE_OUT + i];
    return (int)((o - (S + CVE_OUT)) / 13) + 1;
}
