// Host build of the product's correctly-rounded sin/cos (xuanpolicy_b200/csrc/crtrig.cuh) so the CPU test
// suite can check the algorithm against mpmath without a GPU.  Test infrastructure only.
// Built twice by tests/test_trig_fast_path.py: as is (two-phase) and with -DXB_TRIG_FAST=0 (full series only).
#include "../xuanpolicy_b200/csrc/crtrig.cuh"

extern "C" long host_sincos(const double* x, double* s, double* c, long n) {
    int slow = 0;
    for (long i = 0; i < n; ++i) xb::sincos_cr(x[i], &s[i], &c[i], &slow);
    return slow;
}
