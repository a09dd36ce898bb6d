// examples/shim_phase_demo.cxx -- phase-space neighbours through the reference's interface: a TPHS tree built with
// Aniso = -1 answers FindNearest(tt) with the plain 6D search, the same rows as FindNearestPhase(tt) (reference
// KDFindNearest.cxx:260-262,347-361; PhaseDistSqd, DistFunc.h:41-49).  Checked against a brute-force scan on the host;
// exit code 0 on success.
//   g++ -O2 -std=c++17 -Inbodylib_b200/shim examples/shim_phase_demo.cxx -Lnbodylib_b200 -lnbk -Wl,-rpath,$PWD/nbodylib_b200 -o shim_phase_demo
#include <KDTree.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>
#include <stdexcept>
#include <utility>
#include <vector>
using namespace NBody;

int main() {
    const Int_t M = 20000, K = 8;
    std::vector<Particle> ph(M);
    std::mt19937_64 rng(99);
    std::uniform_real_distribution<double> U(0, 1);
    std::normal_distribution<double> G(0, 1);
    for (Int_t i = 0; i < M; i++) {
        // clumps in position, a velocity spread of a few per cent of the box: both halves of the 6D distance matter
        const double cx = ((i % 50) + 0.5) / 50.0;
        ph[i] = Particle(1.0, std::fmod(cx + 0.01 * G(rng) + 1, 1.0), std::fmod(0.5 + 0.3 * std::sin(i % 50) + 0.01 * G(rng) + 1, 1.0), U(rng),
                         0.02 * G(rng), 0.02 * G(rng), 0.02 * G(rng), i);
        ph[i].SetPID(i);
    }
    int bad = 0;
    for (int periodic = 0; periodic < 2; periodic++) {
        Double_t period[3] = {1, 1, 1};
        KDTree tree(ph.data(), M, 16, KDTree::TPHS, KDTree::KEPAN, 1000, 0, -1, 0, periodic ? period : NULL);
        long wrong = 0;
        for (Int_t tt = 3; tt < M; tt += M / 9) {
            Int_t nn[K], nn2[K], nnx[K];
            Double_t d2[K], d22[K], d2x[K];
            tree.FindNearest(tt, nn, d2, K);
            tree.FindNearestPhase(tt, nn2, d22, K);
            std::vector<std::pair<double, Int_t>> all;
            for (Int_t i = 0; i < M; i++) {
                if (i == tt) continue;
                double best = 1e300;
                for (int im = 0; im < (periodic ? 8 : 1); im++) {           // the reference reflects the QUERY position (DistFunc.h:326-355)
                    double sum = 0;
                    for (int k = 0; k < 6; k++) {
                        double q = ph[tt].GetPhase(k);
                        if (k < 3 && (im >> k & 1)) q = (q < 0.5) ? q + 1.0 : q - 1.0;
                        const double d = q - ph[i].GetPhase(k);
                        sum += d * d;
                    }
                    best = std::min(best, sum);
                }
                if (best > 0) all.push_back(std::make_pair(best, i));
            }
            std::partial_sort(all.begin(), all.begin() + K, all.end());
            for (int j = 0; j < K; j++) wrong += nn[j] != all[j].second || d2[j] != all[j].first || nn2[j] != nn[j] || d22[j] != d2[j];
            Double_t x[6];
            for (int k = 0; k < 6; k++) x[k] = ph[tt].GetPhase(k);
            tree.FindNearest(x, nnx, d2x, K);                       // coordinate form: the particle sitting on x comes first
            wrong += nnx[0] != tt || d2x[0] != 0 || nnx[1] != nn[0] || d2x[1] != d2[0];
        }
        printf("phase-space FindNearest on a %s TPHS tree (Aniso = -1): %ld mismatches against brute force\n", periodic ? "periodic" : "non periodic", wrong);
        if (wrong) bad++;
    }
    // the constructor default Aniso = 0 selects the reference's metric search (quirk Q4): refused, not approximated
    {
        KDTree tree(ph.data(), M, 16, KDTree::TPHS);
        Int_t nn[K]; Double_t d2[K];
        bool threw = false;
        try { tree.FindNearest(0, nn, d2, K); } catch (const std::runtime_error&) { threw = true; }
        if (!threw) bad++;
        tree.FindNearestPhase(0, nn, d2, K);                        // FindNearestPhase itself never looks at Aniso (KDFindNearest.cxx:347-361)
    }
    // velocity-space neighbours on a TVEL tree: FindNearestVel(tt) = FindNearest(tt) = a brute-force scan of the velocities
    {
        KDTree tree(ph.data(), M, 16, KDTree::TVEL);
        long wrong = 0;
        for (Int_t tt = 5; tt < M; tt += M / 9) {
            Int_t nn[K], nn2[K], nnv[K];
            Double_t d2[K], d22[K], d2v[K];
            tree.FindNearestVel(tt, nn, d2, K);
            tree.FindNearest(tt, nn2, d22, K);
            std::vector<std::pair<double, Int_t>> all;
            for (Int_t i = 0; i < M; i++) {
                double sum = 0;
                for (int k = 3; k < 6; k++) { const double d = ph[tt].GetPhase(k) - ph[i].GetPhase(k); sum += d * d; }
                if (i != tt && sum > 0) all.push_back(std::make_pair(sum, i));
            }
            std::partial_sort(all.begin(), all.begin() + K, all.end());
            for (int j = 0; j < K; j++) wrong += nn[j] != all[j].second || d2[j] != all[j].first || nn2[j] != nn[j] || d22[j] != d2[j];
            Double_t v[3] = {ph[tt].GetVelocity(0), ph[tt].GetVelocity(1), ph[tt].GetVelocity(2)};
            tree.FindNearestVel(v, nnv, d2v, K);                     // coordinate form: the particle itself first
            wrong += nnv[0] != tt || d2v[0] != 0 || nnv[1] != nn[0];
        }
        printf("velocity-space FindNearestVel on a TVEL tree: %ld mismatches against brute force\n", wrong);
        if (wrong) bad++;
    }
    {
        // on a position tree the reference would prune velocity queries with position cut planes: refused
        bool threw = false;
        KDTree ptree(ph.data(), M, 16, KDTree::TPHYS);
        Int_t nn[K]; Double_t d2[K];
        try { ptree.FindNearestVel(0, nn, d2, K); } catch (const std::runtime_error&) { threw = true; }
        if (!threw) bad++;
    }
    printf(bad ? "FAILED (%d)\n" : "shim phase demo ok\n", bad);
    return bad ? 1 : 0;
}
