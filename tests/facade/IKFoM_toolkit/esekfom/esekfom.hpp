// TEST MOCK of the reference's esekfom.hpp call surface (the real header needs Eigen + Boost,
// absent here): just enough of esekfom::esekf / state_ikfom for esekfom_b200.hpp to be
// syntax- and type-checked.  Not used by the product.
#pragma once
#include <array>
#include <cstdio>
namespace mock {
struct V3 { double v[3] = {0, 0, 0}; double& operator[](int i) { return v[i]; } double operator[](int i) const { return v[i]; } };
struct Quat { std::array<double, 4> c{{0, 0, 0, 1}}; std::array<double, 4>& coeffs() { return c; } const std::array<double, 4>& coeffs() const { return c; } };
struct S2 { V3 vec; double operator[](int i) const { return vec[i]; } };
struct Cov { double m[23 * 23] = {0}; double& operator()(int i, int j) { return m[i * 23 + j]; } double operator()(int i, int j) const { return m[i * 23 + j]; }
             static Cov Identity() { Cov c; for (int i = 0; i < 23; i++) c(i, i) = 1; return c; } };
}
struct state_ikfom { enum { DOF = 23 }; mock::V3 pos, offset_T_L_I, vel, bg, ba; mock::Quat rot, offset_R_L_I; mock::S2 grav; };
struct input_ikfom {};
namespace esekfom {
template <typename state, int pn, typename input = state, typename measurement = state, int mn = 0>
class esekf {
public:
    typedef mock::Cov cov;
    typedef double scalar_type;
    esekf(const state& x = state(), const cov& P = cov::Identity()) : x_(x), P_(P) {}
    template <class F, class FX, class FW, class H> void init_dyn_share(F, FX, FW, H, int, scalar_type*) {}
    void update_iterated_dyn_share_modified(double, double&) {}
    const state& get_x() const { return x_; }
    const cov& get_P() const { return P_; }
    void change_x(state& s) { x_ = s; }
    void change_P(cov& p) { P_ = p; }
private:
    state x_; cov P_;
};
}
