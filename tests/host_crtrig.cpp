// Host build of the product's correctly-rounded sin/cos (xuanpolicy_b200/csrc/crtrig.cuh) so the CPU test
// suite can check the algorithm against libquadmath/mpmath without a GPU.  Test infrastructure only.
#include "../xuanpolicy_b200/csrc/crtrig.cuh"

extern "C" void host_sincos(const double* x, double* s, double* c, long n) {
    for (long i = 0; i < n; ++i) xb::sincos_cr(x[i], &s[i], &c[i]);
}
