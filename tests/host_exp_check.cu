// TEST-ONLY: exposes the product's exp_glibc (csrc/pmaf_math.cuh) compiled for the host.
#include "../predictive-multi-agent-framework_b200/csrc/pmaf_math.cuh"
extern "C" void hostexp_eval(const double *x, double *y, long n) {
  for (long i = 0; i < n; ++i) y[i] = pmaf::exp_glibc(x[i]);
}
