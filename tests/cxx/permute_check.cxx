// tests/cxx/permute_check.cxx -- CPU check of the shim's host-side particle permutation (nbk_permute_records in
// nbodylib_b200/shim/KDTree.h): byte path (trivially copyable stand-in Particle), move path (a type with a std::string),
// forward and back.  No device call is made.  Compiled and run by tests/test_oracle_cpu.py::test_shim_permutation_on_cpu.
#include <KDTree.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <string>
using namespace NBody;
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct Fat { std::string s; int id; };   // not trivially copyable: takes the move path
int main(int argc, char** argv) {
    long n = argc > 1 ? atol(argv[1]) : 1000003;
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::mt19937_64 rng(1);
    std::shuffle(order.begin(), order.end(), rng);
    std::vector<Particle> parts(n);
    for (long i = 0; i < n; i++) { parts[i] = Particle(1.0 + i, i, 2 * i, 3 * i, 0, 0, 0, (Int_t)i); parts[i].SetPID(i); }
    double t0 = now_s();
    nbk_permute_records(parts.data(), (int64_t)n, [&](int64_t i) { return (int64_t)order[i]; });
    double t1 = now_s();
    int bad = 0;
    for (long i = 0; i < n; i++) bad += parts[i].GetPID() != order[i] || parts[i].X() != order[i] || parts[i].GetMass() != 1.0 + order[i];
    std::vector<int> inv(n);
    for (long i = 0; i < n; i++) inv[parts[i].GetPID()] = (int)i;
    nbk_permute_records(parts.data(), (int64_t)n, [&](int64_t i) { return (int64_t)inv[i]; });
    for (long i = 0; i < n; i++) bad += parts[i].GetPID() != i;
    long m = std::min<long>(n, 50001);
    std::vector<Fat> fat(m);
    for (long i = 0; i < m; i++) { fat[i].s = "particle number " + std::to_string(i) + " with a long enough name to defeat SSO"; fat[i].id = (int)i; }
    std::vector<int> o2(m); std::iota(o2.begin(), o2.end(), 0); std::shuffle(o2.begin(), o2.end(), rng);
    nbk_permute_records(fat.data(), (int64_t)m, [&](int64_t i) { return (int64_t)o2[i]; });
    for (long i = 0; i < m; i++) bad += fat[i].id != o2[i] || fat[i].s.find(" " + std::to_string(o2[i]) + " ") == std::string::npos;
    printf("trivially copyable Particle: %d, n = %ld, permute %.3f s, %d errors\n", (int)std::is_trivially_copyable<Particle>::value, n, t1 - t0, bad);
    return bad ? 1 : 0;
}
