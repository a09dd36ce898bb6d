// sharded_demo.cxx -- the slab-sharded tree from plain C++ (include/nbk_sharded.h), no Python, no torch, no MPI:
// one process per GPU (fork), the 128-byte communicator id travels through a file the way an MPI code would MPI_Bcast it.
//
//   usage: sharded_demo NRANKS [NPARTICLES]
//
// Every rank generates the same clustered particle set, keeps its x-slab, and runs KDTree::CalcDensity and KDTree::FOF over the
// GLOBAL set through nbk_sharded_*; the parent then builds ONE tree over all particles on device 0 (nbk.h) and compares:
// densities to 1e-10, FOF partitions and group counts exactly.  Prints "SHARDED DEMO OK" and exits 0 on success.
//
//   g++ -O2 -std=c++17 -Iinclude examples/sharded_demo.cxx -Lnbodylib_b200 -lnbk_sharded -lnbk -Wl,-rpath,$PWD/nbodylib_b200 -o sharded_demo
#include <nbk_sharded.h>

#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

static uint64_t rng_state;
static double uni() {      // xorshift64*, the same stream in every process
    rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
    return (double)((rng_state * 2685821657736338717ull) >> 11) / 9007199254740992.0;
}
static double gauss() { return std::sqrt(-2.0 * std::log(uni() + 1e-300)) * std::cos(6.283185307179586 * uni()); }

// 60 % uniform background + 40 % in Gaussian blobs (some of them on slab faces), unit periodic box, fp64 coordinates
static void make_particles(int64_t n, std::vector<double>& pos, std::vector<double>& mass) {
    rng_state = 88172645463325252ull;
    pos.resize(3 * (size_t)n); mass.resize((size_t)n);
    const int nblob = 24;
    double c[nblob][3], s[nblob];
    for (int b = 0; b < nblob; b++) { for (int d = 0; d < 3; d++) c[b][d] = uni(); s[b] = 0.004 + 0.01 * uni(); if (b % 4 == 0) c[b][0] = 0.25 * (b / 4 % 4); }
    for (int64_t i = 0; i < n; i++) {
        if (uni() < 0.6) for (int d = 0; d < 3; d++) pos[3 * i + d] = uni();
        else {
            const int b = (int)(uni() * nblob) % nblob;
            for (int d = 0; d < 3; d++) { double x = c[b][d] + s[b] * gauss(); x -= std::floor(x); pos[3 * i + d] = x < 1.0 ? x : 0.0; }
        }
        mass[i] = 0.5 + uni();
    }
}

#define CHECK(call) do { int _rc = (call); if (_rc != NBK_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, _rc, nbk_last_error()); exit(2); } } while (0)

static int run_rank(int rank, int nranks, int64_t n, const std::string& tag, double ll, int k, int minnum) {
    std::vector<double> pos, mass;
    make_particles(n, pos, mass);
    std::vector<double> mp, mm;
    std::vector<int64_t> ids;
    for (int64_t i = 0; i < n; i++) {
        int slab = (int)(pos[3 * i] * nranks); if (slab >= nranks) slab = nranks - 1;
        if (slab == rank) { for (int d = 0; d < 3; d++) mp.push_back(pos[3 * i + d]); mm.push_back(mass[i]); ids.push_back(i); }
    }
    // rendezvous: rank 0 publishes the id (write to a temporary name, then rename: readers never see a partial file)
    unsigned char id[128];
    const std::string idfile = tag + ".id";
    if (rank == 0) {
        CHECK(nbk_comm_unique_id(id));
        FILE* f = fopen((idfile + ".tmp").c_str(), "wb"); fwrite(id, 1, 128, f); fclose(f);
        rename((idfile + ".tmp").c_str(), idfile.c_str());
    } else {
        for (int tries = 0;; tries++) {
            FILE* f = fopen(idfile.c_str(), "rb");
            if (f) { size_t got = fread(id, 1, 128, f); fclose(f); if (got == 128) break; }
            if (tries > 6000) { fprintf(stderr, "rank %d: no communicator id\n", rank); return 3; }
            usleep(10000);
        }
    }
    nbk_comm* comm = nullptr;
    CHECK(nbk_comm_init_rank(nranks, rank, id, rank, &comm));
    nbk_particles p;
    memset(&p, 0, sizeof(p));
    p.pos = mp.data(); p.pos_stride = 24; p.mass = mm.data(); p.mass_stride = 8; p.real_bytes = 8; p.on_device = 0;
    const double box[3] = {1.0, 1.0, 1.0};
    nbk_sharded* st = nullptr;
    CHECK(nbk_sharded_create(comm, &p, (int64_t)mm.size(), box, NULL, /*periodic*/ 1, k, 0.0, &st));
    std::vector<double> rho(mm.size());
    CHECK(nbk_sharded_calc_density(st, k, rho.data(), 0));
    std::vector<int32_t> grp(mm.size());
    int64_t ng = 0;
    CHECK(nbk_sharded_fof(st, -1, ll, NULL, minnum, 1, grp.data(), &ng, 0));
    nbk_sharded_info info;
    CHECK(nbk_sharded_get_info(st, &info));
    printf("rank %d of %d: %lld particles, %lld density ghosts, %lld FOF ghosts, h = %.4g, %lld groups\n", rank, nranks, (long long)info.n_local,
           (long long)info.ghosts_knn, (long long)info.ghosts_fof, info.h_knn, (long long)ng);
    FILE* f = fopen((tag + ".r" + std::to_string(rank)).c_str(), "wb");
    const int64_t m = (int64_t)ids.size();
    fwrite(&m, 8, 1, f); fwrite(&ng, 8, 1, f); fwrite(ids.data(), 8, m, f); fwrite(rho.data(), 8, m, f); fwrite(grp.data(), 4, m, f);
    fclose(f);
    CHECK(nbk_sharded_destroy(st));
    CHECK(nbk_comm_destroy(comm));
    return 0;
}

int main(int argc, char** argv) {
    const int nranks = argc > 1 ? atoi(argv[1]) : 2;
    const int64_t n = argc > 2 ? atoll(argv[2]) : 300000;
    const int k = 32, minnum = 10;
    const double ll = 0.2 / std::cbrt((double)n);
    const std::string tag = "/tmp/nbk_sharded_demo_" + std::to_string((long long)getpid());
    // fork BEFORE the first CUDA call: every child creates its own context on its own device
    std::vector<pid_t> kids;
    for (int r = 0; r < nranks; r++) {
        pid_t pid = fork();
        if (pid == 0) _exit(run_rank(r, nranks, n, tag, ll, k, minnum));
        kids.push_back(pid);
    }
    int bad = 0;
    for (pid_t pid : kids) { int status = 0; waitpid(pid, &status, 0); if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) bad++; }
    if (bad) { fprintf(stderr, "%d rank(s) failed\n", bad); return 1; }
    // gather the ranks' results by global particle index
    std::vector<double> rho(n, -1.0);
    std::vector<int32_t> grp(n, -1);
    int64_t ng_sharded = -1;
    for (int r = 0; r < nranks; r++) {
        const std::string fn = tag + ".r" + std::to_string(r);
        FILE* f = fopen(fn.c_str(), "rb");
        int64_t m = 0, ng = 0;
        if (!f || fread(&m, 8, 1, f) != 1 || fread(&ng, 8, 1, f) != 1) { fprintf(stderr, "cannot read %s\n", fn.c_str()); return 1; }
        std::vector<int64_t> ids(m); std::vector<double> rr(m); std::vector<int32_t> gg(m);
        if (fread(ids.data(), 8, m, f) != (size_t)m || fread(rr.data(), 8, m, f) != (size_t)m || fread(gg.data(), 4, m, f) != (size_t)m) return 1;
        fclose(f); remove(fn.c_str());
        if (ng_sharded >= 0 && ng != ng_sharded) { fprintf(stderr, "ranks disagree on the number of groups\n"); return 1; }
        ng_sharded = ng;
        for (int64_t i = 0; i < m; i++) { rho[ids[i]] = rr[i]; grp[ids[i]] = gg[i]; }
    }
    remove((tag + ".id").c_str());
    // one tree over everything on device 0
    std::vector<double> pos, mass;
    make_particles(n, pos, mass);
    nbk_particles p;
    memset(&p, 0, sizeof(p));
    p.pos = pos.data(); p.pos_stride = 24; p.mass = mass.data(); p.mass_stride = 8; p.real_bytes = 8; p.on_device = 0;
    const double period[3] = {1.0, 1.0, 1.0};
    nbk_tree* t = nullptr;
    CHECK(nbk_create(&p, n, 16, NBK_TPHYS, NBK_KEPAN, 1000, 0, period, 0, 0, &t));
    std::vector<double> rho1(n);
    CHECK(nbk_calc_density(t, k, rho1.data(), NULL, 0));
    std::vector<int32_t> grp1(n);
    int64_t ng1 = 0;
    CHECK(nbk_fof(t, ll, minnum, 1, NULL, grp1.data(), &ng1, NULL, 0));
    CHECK(nbk_destroy(t));
    double worst = 0;
    int64_t unset = 0, mismatched = 0;
    std::map<int32_t, int32_t> fwd, bwd;
    for (int64_t i = 0; i < n; i++) {
        if (rho[i] < 0 || grp[i] < 0) { unset++; continue; }
        worst = std::fmax(worst, std::fabs(rho[i] - rho1[i]) / rho1[i]);
        if ((grp[i] == 0) != (grp1[i] == 0)) { mismatched++; continue; }
        if (grp[i] == 0) continue;
        auto a = fwd.emplace(grp[i], grp1[i]); auto b = bwd.emplace(grp1[i], grp[i]);
        if (a.first->second != grp1[i] || b.first->second != grp[i]) mismatched++;
    }
    printf("%d ranks, %lld particles: density worst relative difference %.3g, %lld groups (single tree: %lld), %lld particles in a different group, %lld unset\n",
           nranks, (long long)n, worst, (long long)ng_sharded, (long long)ng1, (long long)mismatched, (long long)unset);
    if (unset || mismatched || ng_sharded != ng1 || !(worst < 1e-10)) { printf("SHARDED DEMO FAILED\n"); return 1; }
    printf("SHARDED DEMO OK\n");
    return 0;
}
