// examples/shim_bench.cxx -- the drop-in path timed end to end: an array of 88-byte NBody::Particle records (fp64 fields, the
// reference's default layout) goes through the header-only shim exactly as VELOCIraptor would drive it:
//     KDTree tree(parts, N, 16, TPHYS, KEPAN, 1000, 0, 0, 0, period);   // SetID, strided H2D, build, permute the array into tree order
//     tree.CalcDensity(k);                                              // kNN + SPH density, rho written into the particles
//     ~KDTree                                                           // array restored to input order
// Built by __graft_entry__.build() into nbodylib_b200/libnbk_shimbench.so; bench.py calls nbk_shim_e2e through ctypes and
// reports the result as `e2e_aos`.
#include <KDTree.h>

#include <chrono>
#include <cstdio>
#include <vector>
using namespace NBody;

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

extern "C" int nbk_shim_e2e(const float* pos, const float* vel, const float* mass, long n, int k, const double* period, int reps,
                            double* seconds /* [reps][4]: constructor, CalcDensity, destructor, total */, double* rho_sum, char* err, int errlen) {
    try {
        std::vector<Particle> parts((size_t)n);
        NBK_SHIM_PARALLEL_FOR
        for (long i = 0; i < n; i++) {
            parts[i] = Particle(mass ? mass[i] : 1.0, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], vel ? vel[3 * i] : 0.0, vel ? vel[3 * i + 1] : 0.0,
                                vel ? vel[3 * i + 2] : 0.0, (Int_t)i);
            parts[i].SetPID((Int_t)i);
        }
        Double_t per[3] = {0, 0, 0};
        if (period) for (int j = 0; j < 3; j++) per[j] = period[j];
        for (int r = 0; r < reps; r++) {
            const double t0 = now_s();
            double t1, t2;
            {
                KDTree tree(parts.data(), (Int_t)n, 16, KDTree::TPHYS, KDTree::KEPAN, 1000, 0, 0, 0, period ? per : NULL);
                t1 = now_s();
                tree.CalcDensity(k);
                t2 = now_s();
            }
            const double t3 = now_s();
            seconds[4 * r] = t1 - t0; seconds[4 * r + 1] = t2 - t1; seconds[4 * r + 2] = t3 - t2; seconds[4 * r + 3] = t3 - t0;
        }
        double s = 0;
        bool order_ok = true;
        for (long i = 0; i < n; i++) { s += parts[i].GetDensity(); order_ok = order_ok && parts[i].GetID() == (Int_t)i && parts[i].GetPID() == (Int_t)i; }
        *rho_sum = s;
        if (!order_ok) { std::snprintf(err, errlen, "particle order was not restored"); return 2; }
        return 0;
    } catch (const std::exception& e) {
        std::snprintf(err, errlen, "%s", e.what());
        return 1;
    }
}
