// Host build of csrc/special.cuh for tests/test_host_logic.py (g++ -shared): the same source the device compiles.
#include "../nabla.jl_b200/csrc/special.cuh"
extern "C" {
double sf_airy(double x, int which) { return nb_airy(x, which); }
double sf_dawson(double x) { return nb_dawson(x); }
double sf_erfi(double x) { return nb_erfi(x); }
double sf_invdigamma(double y) { return nb_invdigamma(y); }
double sf_digamma_pos(double x) { return nb_sf_digamma_pos(x); }
double sf_trigamma_pos(double x) { return nb_sf_trigamma_pos(x); }
}
extern "C" {
double sf_bessel(int which, double nu, double x) { return nb_bessel(which, nu, x); }
double sf_bessel_dx(int which, double nu, double x) { return nb_bessel_dx(which, nu, x); }
double sf_polygamma(double n, double x) { return nb_polygamma_n(n, x); }
}
