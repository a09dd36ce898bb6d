/* oracle/harness_main.cxx -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the reference's own test harness, src/tests/test_kdtree.cxx, compiled UNCHANGED (its main() renamed on the
 * compiler command line) with the reference's src/NBody and src/Math headers and nbodylib_b200/shim/KDTree.h in place of
 * the reference's src/KDTree/KDTree.h: the proof that a consumer of NBody::KDTree switches to the B200 library by changing
 * its include path and link line.  oracle/Makefile target `harness` -> oracle/_ref/test_kdtree_shim
 * (built where /root/reference exists; the binary travels to the GPU box).
 *
 * The harness functions are called for its five tree types (test_kdtree.cxx:48-58) where the calls have device implementations:
 *   Physical                (TPHYS, b = 16): kdtree_test_NN, kdtree_test_ballsearch, kdtree_test_FOF
 *   Physical Rdist          (TPHYS, b = 0.01 N, Rdistadapt = 0.01): the same three
 *   Physical Rdist Adaptfac (TPHYS, b = 0.01 N, Rdistadapt = 0.01, AdaptiveMedianFac = 0.1): the same three
 *   Velocity                (TVEL,  b = 16): kdtree_test_NN
 *   Phase                   (TPHS,  b = 16): kdtree_test_ballsearch, kdtree_test_FOF   (its FindNearest takes the reference's metric path: Aniso = 0)
 * followed by brute-force checks of what the harness only prints.  The Rdist options shape the reference's tree only; the
 * shim serves the same results from the device tree (shim/KDTree.h constructor).
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <NBodyMath.h>
#include <NBody.h>
#include <KDTree.h>

// defined in the reference's src/tests/test_kdtree.cxx
std::vector<NBody::Particle> generate_vector(std::size_t size, double fac, int nsub);
NBody::KDTree* build_kdtree(std::vector<NBody::Particle>& parts, Int_t b, int treetype, double rdist2fac, double adaptivefactor, double bfac);
void kdtree_test_NN(NBody::KDTree*& tree, std::vector<NBody::Particle>& parts, int num_nn);
void kdtree_test_ballsearch(NBody::KDTree*& tree, std::vector<NBody::Particle>& parts, double rdist);
void kdtree_test_FOF(NBody::KDTree*& tree, std::vector<NBody::Particle>& parts, Int_t minnum, Double_t rdist);

static int failures = 0;
#define EXPECT(cond, what) do { if (!(cond)) { failures++; std::printf("CHECK FAILED: %s\n", what); } } while (0)

static double d2(const NBody::Particle& a, const NBody::Particle& b, int off) {
    double t = 0;
    for (int j = 0; j < 3; j++) { double d = a.GetPhase(off + j) - b.GetPhase(off + j); t += d * d; }
    return t;
}

int main(int argc, char** argv) {
    const std::size_t N = argc > 1 ? (std::size_t)std::atol(argv[1]) : 20000;
    std::vector<NBody::Particle> parts = generate_vector(N, 0.1, 100);
    std::vector<NBody::Particle> input(parts);
    const int k = 16;
    const char* names[5] = {"Physical", "Velocity", "Phase", "Physical Rdist", "Physical Rdist Adaptfac"};
    for (int cfg = 0; cfg < 5; cfg++) {
        const int which = cfg < 3 ? cfg : 0;             // 0: position tree, 1: velocity tree, 2: phase-space tree
        const int tt = which == 0 ? NBody::KDTree::TPHYS : (which == 1 ? NBody::KDTree::TVEL : NBody::KDTree::TPHS);
        std::printf("==== %s tree\n", names[cfg]);
        // arguments of the reference's TreeTypes() table (test_kdtree.cxx:48-58)
        NBody::KDTree* tree = cfg < 3 ? build_kdtree(parts, 16, tt, -1, 0.0, 0.0) : build_kdtree(parts, 0, tt, 0.01, cfg == 4 ? 0.1 : 0.0, 0.01);
        if (cfg >= 3) EXPECT(tree->GetBucketSize() == (Int_t)(0.01 * N), "GetBucketSize() returns the caller's leaf size");
        EXPECT(tree->GetNumLeafNodes() > 0 && tree->GetNumNodes() == 2 * tree->GetNumLeafNodes() - 1, "node counts of a binary tree");
        if (which != 2) {
            kdtree_test_NN(tree, parts, k);
            // the reference's harness prints statistics only: check a sample against brute force (same fp64 expression)
            std::vector<Int_t> nn(k); std::vector<Double_t> nd(k);
            for (std::size_t i = 0; i < N; i += N / 200 + 1) {
                tree->FindNearest((Int_t)i, nn.data(), nd.data(), k);
                std::vector<double> all;
                for (std::size_t j = 0; j < N; j++) { double d = d2(parts[i], parts[j], which == 1 ? 3 : 0); if (j != i && d > 0) all.push_back(d); }
                std::partial_sort(all.begin(), all.begin() + k, all.end());
                bool same = true;
                for (int j = 0; j < k; j++) same = same && all[j] == nd[j] && d2(parts[i], parts[nn[j]], which == 1 ? 3 : 0) == nd[j];
                EXPECT(same, "FindNearest(i) == brute force");
            }
        }
        if (which != 1) {
            kdtree_test_ballsearch(tree, parts, 0.1);
            for (std::size_t i = 0; i < N; i += N / 100 + 1) {
                std::vector<Int_t> v = tree->SearchBallPosTagged((Int_t)i, 0.01);
                std::size_t c = 0;
                for (std::size_t j = 0; j < N; j++) if (j != i && d2(parts[i], parts[j], 0) < 0.01) c++;
                EXPECT(v.size() == c, "SearchBallPosTagged(i) count == brute force");
            }
            kdtree_test_FOF(tree, parts, 20, 1.0);
            Int_t ng = 0;
            Int_t* g = tree->FOF(1.0, ng, 20);
            bool one = ng == 1;
            for (std::size_t i = 0; one && i < N; i++) one = g[i] == 1;
            EXPECT(one, "FOF with a box-sized linking length: one group holding every particle");
            delete[] g;
        }
        delete tree;
        bool restored = true;
        for (std::size_t i = 0; restored && i < N; i++) restored = parts[i].GetID() == (Int_t)i && parts[i].X() == input[i].X() && parts[i].GetVelocity(2) == input[i].GetVelocity(2);
        EXPECT(restored, "~KDTree restores the caller's particle order");
    }
    std::printf(failures ? "HARNESS FAILED (%d checks)\n" : "HARNESS OK\n", failures);
    return failures ? 1 : 0;
}
