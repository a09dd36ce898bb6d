// examples/shim_demo.cxx -- a program written against the reference's NBody::KDTree interface (it mirrors what
// reference src/tests/test_kdtree.cxx does: build, per-particle FindNearest, SearchBallPosTagged, FOF), compiled
// against the shim header and linked with libnbk.so.  Prints a few invariants; exit code 0 on success.
//   g++ -O2 -std=c++17 -Inbodylib_b200/shim examples/shim_demo.cxx -Lnbodylib_b200 -lnbk -Wl,-rpath,$PWD/nbodylib_b200 -o shim_demo
#include <KDTree.h>

#include <chrono>
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <random>
using namespace NBody;

static int type_check(Particle& p, Double_t*) { return p.GetType() != 0 ? -1 : 0; }

int main() {
    const Int_t N = 200000;
    std::vector<Particle> parts(N);
    std::mt19937_64 rng(4322);
    std::uniform_real_distribution<double> U(0, 1);
    std::normal_distribution<double> G(0, 1);
    for (Int_t i = 0; i < N; i++) {
        double c[3];
        if (i % 10 == 0) { c[0] = U(rng); c[1] = U(rng); c[2] = U(rng); }
        else { double cx = ((i % 100) + 0.5) / 100.0; c[0] = std::fmod(cx + 0.002 * G(rng) + 1, 1.0); c[1] = std::fmod(0.5 + 0.3 * std::sin(i % 100) + 0.002 * G(rng) + 1, 1.0); c[2] = std::fmod(0.5 + 0.002 * G(rng) + 1, 1.0); }
        parts[i] = Particle(1.0, (float)c[0], (float)c[1], (float)c[2], (float)G(rng), (float)G(rng), (float)G(rng), i);
        parts[i].SetPID(i);
    }
    Double_t period[3] = {1, 1, 1};
    int bad = 0;
    {
        KDTree tree(parts.data(), N, 16, KDTree::TPHYS, KDTree::KEPAN, 1000, 0, 0, 0, period);
        printf("nodes %d leaves %d kernnorm %.17g\n", tree.GetNumNodes(), tree.GetNumLeafNodes(), tree.GetKernNorm());
        // the array is now in tree order, ids hold the input index
        Int_t nn[16]; Double_t d2[16];
        for (Int_t tt = 0; tt < N; tt += N / 7) {
            tree.FindNearest(tt, nn, d2, 16);
            for (int j = 0; j < 16; j++) {
                double s = 0;
                for (int k = 0; k < 3; k++) { double d = parts[tt].GetPosition(k) - parts[nn[j]].GetPosition(k); d -= std::round(d); s += d * d; }
                if (std::fabs(s - d2[j]) > 1e-12 * (s + 1e-30) || nn[j] == tt || (j && d2[j] < d2[j - 1])) bad++;
            }
        }
        // the reference's own usage pattern (tests/test_kdtree.cxx:279-301): an OpenMP loop over the per-particle call.  The
        // shim serves it from per-thread block caches filled by batched device queries; it must agree with the whole-system form
        {
            const int K = 8;
            std::vector<Int_t> nn_loop((size_t)N * K);
            std::vector<Double_t> d2_loop((size_t)N * K);
#pragma omp parallel for schedule(guided)
            for (Int_t i = 0; i < N; i++) tree.FindNearestPos(i, &nn_loop[(size_t)i * K], &d2_loop[(size_t)i * K], K);
            std::vector<Int_t*> nnp(N); std::vector<Double_t*> d2p(N);
            std::vector<Int_t> nn_all((size_t)N * K); std::vector<Double_t> d2_all((size_t)N * K);
            for (Int_t i = 0; i < N; i++) { nnp[i] = &nn_all[(size_t)i * K]; d2p[i] = &d2_all[(size_t)i * K]; }
            tree.FindNearestPos(nnp.data(), d2p.data(), K);
            long diff = 0;
            for (size_t q = 0; q < nn_all.size(); q++) diff += (nn_all[q] != nn_loop[q]) || (d2_all[q] != d2_loop[q]);
            printf("per-particle loop vs whole system: %ld differences\n", diff);
            if (diff) bad++;
        }
        std::vector<Int_t> tagged = tree.SearchBallPosTagged(N / 2, 0.01 * 0.01);
        printf("ball: %zu particles\n", tagged.size());
        tree.CalcDensity(32);
        double mean = 0;
        for (Int_t i = 0; i < N; i++) mean += parts[i].GetDensity();
        printf("mean density %.6g\n", mean / N);
        Int_t ng = 0;
        Int_t* pfof = tree.FOF(0.2 / std::cbrt((double)N), ng, 20, 1);
        long grouped = 0;
        for (Int_t i = 0; i < N; i++) grouped += pfof[i] > 0;
        printf("FOF: %d groups, %ld grouped\n", ng, grouped);
        delete[] pfof;
        Double_t params[10] = {0};
        params[1] = params[6] = std::pow(0.2 / std::cbrt((double)N), 2); params[2] = params[7] = 1.0;
        Int_t ng6 = 0;
        Int_t* p6 = tree.FOFCriterion(FOF6d, params, ng6, 20);
        printf("FOF6d: %d groups\n", ng6);
        delete[] p6;
        if (ng <= 0 || !(mean > 0)) bad++;
    }
    // second tree, non periodic: the single-target estimators, criterion search, dense ball search and the node mirror
    {
        KDTree tree(parts.data(), N, 16, KDTree::TPHYS, KDTree::KEPAN, 1000);
        const int K = 32;
        Int_t nn[K]; Double_t d2[K], dist[K], weight[K];
        double worst = 0;
        for (Int_t tt = 3; tt < N; tt += N / 11) {
            // CalcDensityParticle == CalcSmoothLocalValue over the target's own neighbour list with the masses as weights
            tree.FindNearestPos(tt, nn, d2, K);
            for (int j = 0; j < K; j++) { dist[j] = std::sqrt(d2[K - 1 - j]); weight[j] = parts[nn[K - 1 - j]].GetMass(); }
            double v1 = tree.CalcSmoothLocalValue(K, dist, weight), v2 = tree.CalcDensityParticle(tt, K);
            PriorityQueue pq(K);
            for (int j = 0; j < K; j++) pq.Push(nn[j], d2[j]);
            for (int j = 0; j < K; j++) weight[j] = 1.0;                       // unit masses: the pop order does not matter
            double v3 = tree.CalcSmoothLocalValue(K, &pq, weight);
            worst = std::fmax(worst, std::fmax(std::fabs(v1 - v2), std::fabs(v3 - v2)) / v2);
            // the position form at the particle's own position sees the particle itself as the nearest neighbour
            Double_t x[3] = {parts[tt].X(), parts[tt].Y(), parts[tt].Z()}, v[3] = {parts[tt].GetVelocity(0), parts[tt].GetVelocity(1), parts[tt].GetVelocity(2)};
            if (!(tree.CalcDensityPosition(x, K) > 0) || !(tree.CalcVelDensityPosition(x, v, K / 2, K) > 0) || !(tree.CalcVelDensityParticle(tt, K / 2, K) > 0)) bad++;
            // criterion search with FOF3d == ball search of the same radius; dense form marks the same particles
            const double r2 = 0.01 * 0.01;
            Double_t params[10] = {0};
            params[1] = params[6] = r2;
            std::vector<Int_t> a = tree.SearchBallPosTagged(tt, r2), b = tree.SearchCriterionTagged(tt, FOF3d, params);
            std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
            if (a != b) bad++;
            std::vector<Int_t> mark(N, 0); std::vector<Double_t> md2(N, -1.0);
            tree.SearchBallPos(tt, r2, 7, mark.data(), md2.data());
            size_t marked = 0;
            for (Int_t i = 0; i < N; i++) if (mark[i] == 7) { marked++; if (!(md2[i] >= 0 && md2[i] < r2)) bad++; }
            if (marked != a.size()) bad++;
            for (Int_t j : a) if (mark[parts[j].GetID()] != 7) bad++;
            // node mirror
            Node* leaf = tree.FindLeafNode(tt);
            if (!(leaf->GetLeaf() && leaf->GetStart() <= tt && tt < leaf->GetEnd() && leaf->GetCount() <= 16)) bad++;
            for (int k = 0; k < 3; k++) if (!(leaf->GetBoundary(k, 0) <= x[k] && x[k] <= leaf->GetBoundary(k, 1))) bad++;
            if (tree.FindLeafNode(x) != leaf && tree.FindLeafNode(x)->GetCount() > 16) bad++;
        }
        // filtered neighbours: FindNearestCheck only returns particles passing the caller's check, FindNearestCriterion only
        // particles meeting the criterion; both are subsets of the unfiltered ordering
        {
            for (Int_t i = 0; i < N; i++) parts[i].SetType(parts[i].GetPID() % 3 == 0 ? 1 : 0);
            Double_t params[10] = {0};
            params[1] = params[6] = 0.02 * 0.02; params[2] = params[7] = 4.0;
            Int_t n8[8], n8c[8]; Double_t d8[8], d8c[8], dall[8]; Int_t nall[8];
            for (Int_t tt = 5; tt < N; tt += N / 13) {
                tree.FindNearestCheck(tt, type_check, params, n8, d8, 8);
                tree.FindNearestPos(tt, nall, dall, 8);
                for (int j = 0; j < 8; j++) {
                    if (n8[j] < 0 || parts[n8[j]].GetType() != 0 || n8[j] == tt || (j && d8[j] < d8[j - 1]) || d8[j] < dall[j]) bad++;
                }
                tree.FindNearestCriterion(tt, FOF6d, params, n8c, d8c, 8);
                for (int j = 0; j < 8; j++) {
                    if (n8c[j] < 0) { if (d8c[j] < 1e31) bad++; continue; }
                    if (!FOF6d(parts[tt], parts[n8c[j]], params) || n8c[j] == tt || (j && d8c[j] < d8c[j - 1])) bad++;
                }
            }
            for (Int_t i = 0; i < N; i++) parts[i].SetType(0);
        }
        // smoothed velocity field: with densities set, the mean velocity of a particle is a weighted mean of its neighbours'
        {
            tree.CalcDensity(32);
            Coordinate* sv = tree.CalcSmoothVel(32);
            Matrix* sd = tree.CalcSmoothVelDisp(sv, 32);
            double tr = 0;
            for (Int_t i = 0; i < N; i += 997) {
                for (int j = 0; j < 3; j++) if (!(std::fabs(sv[i][j]) < 50.0) || !(sd[i](j, j) >= 0)) bad++;
                if (std::fabs(sd[i](0, 1) - sd[i](1, 0)) > 1e-9 * (sd[i](0, 0) + sd[i](1, 1) + 1e-30)) bad++;
                tr += sd[i](0, 0) + sd[i](1, 1) + sd[i](2, 2);
            }
            printf("smoothed velocity dispersion: mean trace %.4g\n", tr / ((N + 996) / 997));
            if (!(tr > 0)) bad++;
            // higher moments: finite, and the reference's kurtosis form carries -3 per contribution (about 2 x 32 of them)
            Coordinate* sk = tree.CalcSmoothVelSkew(sv, sd, 32);
            Coordinate* ku = tree.CalcSmoothVelKurtosis(sv, sd, 32);
            double kmean = 0;
            for (Int_t i = 0; i < N; i += 997) {
                for (int j = 0; j < 3; j++) if (!std::isfinite(sk[i][j]) || !std::isfinite(ku[i][j])) bad++;
                kmean += ku[i][0];
            }
            kmean /= ((N + 996) / 997);
            printf("smoothed velocity kurtosis (reference form): mean %.4g\n", kmean);
            if (!(kmean < -100.0 && kmean > -300.0)) bad++;
            delete[] sk; delete[] ku;
            delete[] sv; delete[] sd;
        }
        printf("single-target density vs CalcSmoothLocalValue: worst relative difference %.3g\n", worst);
        if (!(worst < 1e-12)) bad++;
        long leaves = 0, count = 0;
        std::vector<Node*> st(1, tree.GetRoot());
        while (!st.empty()) {
            Node* nd = st.back(); st.pop_back();
            if (nd->GetLeaf()) { leaves++; count += nd->GetCount(); }
            else { st.push_back(((SplitNode*)nd)->GetRight()); st.push_back(((SplitNode*)nd)->GetLeft()); }
        }
        printf("node mirror: %ld leaves, %ld particles\n", leaves, count);
        if (leaves != tree.GetNumLeafNodes() || count != N || tree.GetRoot()->GetCount() != N) bad++;
    }
    // destructor restored the input order
    for (Int_t i = 0; i < N; i++) if (parts[i].GetID() != i || parts[i].GetPID() != i) { bad++; break; }
    // third tree: built from a System (KDTree.cxx:1322-1338: the period comes from the System, all-zero = open box), the
    // per-particle ball search loop of reference tests/test_kdtree.cxx:325-341 against the batched C-ABI call, growing one group
    // from a particle (FOFCriterionParticle), and OverWriteInputOrder (the array stays in tree order with fresh ids)
    {
        System S(N, parts.data(), Coordinate(1.0, 1.0, 1.0));
        KDTree tree(S, 16, KDTree::TPHYS, KDTree::KEPAN, 1000);
        if (tree.GetPeriod(0) != 1.0 || tree.GetNumLeafNodes() <= 0) bad++;
        const double r2 = 0.004 * 0.004;
        std::vector<int> cnt_loop(N);
        auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(guided)
        for (Int_t i = 0; i < N; i++) cnt_loop[i] = (int)tree.SearchBallPosTagged(i, r2).size();
        const double t_loop = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::vector<int32_t> q(N), idx;
        std::vector<int64_t> off((size_t)N + 1);
        for (Int_t i = 0; i < N; i++) q[i] = i;
        int64_t tot = 0;
        t0 = std::chrono::steady_clock::now();
        nbk_ball_particles(tree.GetHandle(), r2, N, q.data(), off.data(), NULL, NULL, 0, &tot, 0);
        idx.resize((size_t)std::max<int64_t>(tot, 1));
        if (nbk_ball_particles(tree.GetHandle(), r2, N, q.data(), off.data(), idx.data(), NULL, (int64_t)idx.size(), &tot, 0) != NBK_OK) bad++;
        const double t_batch = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        long diff = 0;
        for (Int_t i = 0; i < N; i++) diff += cnt_loop[i] != (int)(off[(size_t)i + 1] - off[i]);
        printf("SearchBallPosTagged(i) loop over %d particles: %.3f s; batched nbk_ball_particles: %.3f s (ratio %.2f); %ld rows differ\n", N, t_loop, t_batch,
               t_loop / t_batch, diff);
        if (diff) bad++;
        // grow group 7 from one particle with FOF3d: the same set as the FOF group of that particle
        const double ll = 0.2 / std::cbrt((double)N);
        Int_t ng = 0;
        Int_t* pfof = tree.FOF(ll, ng, 2, 0);
        Int_t seed = -1;
        for (Int_t i = 0; i < N && seed < 0; i++) if (pfof[parts[i].GetID()] > 0) seed = i;
        if (seed >= 0) {
            Double_t params[10] = {0};
            params[1] = params[6] = ll * ll;
            std::vector<Int_t> tags(N, 0);
            std::vector<Int_tree_t> plen(16, 0);
            Int_t sz = tree.FOFCriterionParticle(FOF3d, tags.data(), seed, 7, params, NULL, NULL, NULL, NULL, NULL, plen.data());
            const Int_t gseed = pfof[parts[seed].GetID()];
            long mism = 0, members = 0;
            for (Int_t i = 0; i < N; i++) { members += pfof[i] == gseed; mism += (pfof[i] == gseed) != (tags[i] == 7); }
            printf("FOFCriterionParticle: group of %d, FOF group of the seed %ld, %ld mismatches\n", sz, members, mism);
            if (mism || sz != members || plen[7] != sz) bad++;
        } else bad++;
        delete[] pfof;
        tree.OverWriteInputOrder();
    }
    // OverWriteInputOrder: the destructor left the array in tree order and the ids number that order
    {
        bool ids = true, moved = false;
        for (Int_t i = 0; i < N; i++) { ids = ids && parts[i].GetID() == i; moved = moved || parts[i].GetPID() != i; }
        if (!ids || !moved) bad++;
    }
    printf(bad ? "FAILED (%d)\n" : "shim demo ok\n", bad);
    return bad ? 1 : 0;
}
