/*! \file nbk_standalone_types.h (nbodylib_b200 shim)
 *  Minimal stand-in for the reference's NBody::Particle / NBody::Coordinate / NBody::System (reference
 *  src/NBody/Particle.h:264-354, accessors :452-532; src/Math/Coordinate.h:44-70; src/NBody/System.h:59-81) for
 *  programs that do not have the NBodylib headers.  Same default-build field layout (sizeof == 88: mass@0,
 *  position@8, velocity@32, pid@56, id@60, type@64, rho@72, phi@80) and the accessors the tree uses.
 *  A consumer that has the real headers defines NBK_USE_REFERENCE_PARTICLE: shim/KDTree.h then includes the reference's
 *  <NBody.h>, <NBodyMath.h>, <PriorityQueue.h> and <FOFFunc.h> instead of this file.  (The file is deliberately NOT called
 *  Particle.h: with the shim directory first on the include path it must never shadow the reference's <Particle.h>.)
 */
#ifndef NBK_SHIM_PARTICLE_H
#define NBK_SHIM_PARTICLE_H
#include <cstddef>
#include <queue>
#include <stdexcept>
#include <utility>
#include <vector>

namespace NBody {
typedef double Double_t;
typedef float Real_t;
typedef int Int_t;
typedef unsigned int UInt_t;
typedef Int_t Int_tree_t;
typedef UInt_t UInt_tree_t;

class Coordinate {
    Double_t c[3];
public:
    Coordinate(Double_t x = 0, Double_t y = 0, Double_t z = 0) { c[0] = x; c[1] = y; c[2] = z; }
    Coordinate(const Double_t* p) { c[0] = p[0]; c[1] = p[1]; c[2] = p[2]; }
    Double_t& operator[](int i) { return c[i]; }
    const Double_t& operator[](int i) const { return c[i]; }
    Double_t* GetCoord() { return c; }
    const Double_t* GetCoord() const { return c; }
};

/// 3x3 matrix with the element access of the reference's Math::Matrix (Matrix.h): the return type of CalcSmoothVelDisp
class Matrix {
    Double_t m[3][3];
public:
    Matrix(Double_t a = 0) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = a; }
    Double_t& operator()(int i, int j) { return m[i][j]; }
    const Double_t& operator()(int i, int j) const { return m[i][j]; }
};

class Particle {
protected:
    Double_t mass;
    Double_t position[3];
    Double_t velocity[3];
    Int_t pid;
    Int_t id;
    int type;
    Double_t rho;
    Double_t phi;
public:
    Particle(Double_t Mass = 0, Double_t x = 0, Double_t y = 0, Double_t z = 0, Double_t vx = 0, Double_t vy = 0, Double_t vz = 0,
             Int_t ID = 0, int Type = 0, Double_t Rho = 0, Double_t Phi = 0, Int_t PID = 0)
        : mass(Mass), pid(PID), id(ID), type(Type), rho(Rho), phi(Phi) {
        position[0] = x; position[1] = y; position[2] = z; velocity[0] = vx; velocity[1] = vy; velocity[2] = vz;
    }
    Double_t GetMass() const { return mass; }
    void SetMass(const Double_t& m) { mass = m; }
    const Double_t* GetPosition() const { return position; }
    Double_t GetPosition(const int& i) const { return position[i]; }
    Double_t X() const { return position[0]; }
    Double_t Y() const { return position[1]; }
    Double_t Z() const { return position[2]; }
    void SetPosition(const int& i, const Double_t& x) { position[i] = x; }
    void SetPosition(const Double_t& x, const Double_t& y, const Double_t& z) { position[0] = x; position[1] = y; position[2] = z; }
    const Double_t* GetVelocity() const { return velocity; }
    Double_t GetVelocity(const int& i) const { return velocity[i]; }
    void SetVelocity(const int& i, const Double_t& x) { velocity[i] = x; }
    void SetVelocity(const Double_t& x, const Double_t& y, const Double_t& z) { velocity[0] = x; velocity[1] = y; velocity[2] = z; }
    Double_t GetPhase(const int& i) const { return i < 3 ? position[i] : velocity[i - 3]; }
    Int_t GetPID() const { return pid; }
    void SetPID(const Int_t& i) { pid = i; }
    Int_t GetID() const { return id; }
    void SetID(const Int_t& i) { id = i; }
    int GetType() const { return type; }
    void SetType(int i) { type = i; }
    Double_t GetDensity() const { return rho; }
    void SetDensity(const Double_t& r) { rho = r; }
    Double_t GetPotential() const { return phi; }
    void SetPotential(const Double_t& p) { phi = p; }
    /// reference Particle.h:666
    void ScalePhase(Double_t& x, Double_t& v) {
        position[0] *= x; position[1] *= x; position[2] *= x; velocity[0] *= v; velocity[1] *= v; velocity[2] *= v;
    }
};

class System {
    Int_t numparts;
    Particle* particle;
    Coordinate period;
public:
    System(Int_t n, Particle* p, const Coordinate& per = Coordinate(0, 0, 0)) : numparts(n), particle(p), period(per) {}
    Particle* Parts() { return particle; }
    Int_t GetNumParts() const { return numparts; }
    Coordinate GetPeriod() const { return period; }
};

/// Bounded max-priority queue with the interface of the reference's NBody::PriorityQueue (PriorityQueue.h:14-85: Push / Pop /
/// TopPriority / TopQueue / Size / MaxSize / Reset / Empty): the argument type of CalcSmoothLocalValue and the scratch type of
/// CalcVelDensityParticle.  Built on std::priority_queue; among equal priorities the pop order is unspecified, as it is in
/// the reference (an implicit binary heap).
class PriorityQueue {
    typedef std::pair<Double_t, Int_t> item;
    std::priority_queue<item> q;
    Int_t max_size;
public:
    explicit PriorityQueue(Int_t max) : max_size(max) {}
    bool Empty() const { return q.empty(); }
    Int_t Size() const { return (Int_t)q.size(); }
    Int_t MaxSize() const { return max_size; }
    void Reset() { q = std::priority_queue<item>(); }
    void Push(Int_t p, Double_t dist) {
        if ((Int_t)q.size() >= max_size) throw std::runtime_error("PriorityQueue: pushing beyond the allocated size");   // reference: exit()
        q.push(item(dist, p));
    }
    void Pop() { q.pop(); }
    Double_t TopPriority() const { return q.top().first; }
    Int_t TopQueue() const { return q.top().second; }
};

// FOF criteria of the reference (FOFFunc.h:24-57).  Only their ADDRESSES matter to the device tree: the shim maps
// &FOF3d / &FOF6d to the device enum; the bodies are kept so host code can still call them.
typedef int (*FOFcompfunc)(Particle&, Particle&, Double_t*);
typedef int (*FOFcheckfunc)(Particle&, Double_t*);
inline int FOF3d(Particle& a, Particle& b, Double_t* params) {
    Double_t total = 0;
    for (int j = 0; j < 3; j++) total += (a.GetPosition(j) - b.GetPosition(j)) * (a.GetPosition(j) - b.GetPosition(j)) / params[6];
    return (total < 1);
}
inline int FOFVel(Particle& a, Particle& b, Double_t* params) {
    Double_t total = 0;
    for (int j = 0; j < 3; j++) total += (a.GetVelocity(j) - b.GetVelocity(j)) * (a.GetVelocity(j) - b.GetVelocity(j)) / params[6];
    return (total < 1);
}
inline int FOF6d(Particle& a, Particle& b, Double_t* params) {
    Double_t total = 0;
    for (int j = 0; j < 3; j++) {
        total += (a.GetPosition(j) - b.GetPosition(j)) * (a.GetPosition(j) - b.GetPosition(j)) / params[6];
        total += (a.GetVelocity(j) - b.GetVelocity(j)) * (a.GetVelocity(j) - b.GetVelocity(j)) / params[7];
    }
    return (total < 1);
}
inline int Pnocheck(Particle&, Double_t*) { return 0; }
}  // namespace NBody
#endif
