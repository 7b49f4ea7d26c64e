// Study harness (CPU): which species drive the step-size controller?  Builds the solver core with hooks that (a) record
// each species' share of the squared error norm over accepted+rejected steps, (b) apply per-species weights to it.
//   g++ -O2 -shared -fPIC -std=c++17 -ffp-contract=off -I gync.jl_b200/csrc tools/rodas5/errweights.cpp -o /tmp/errw.so
#include <math.h>
#include <string.h>
static double g_errw[33] = {1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1};
static double g_errshare[33];
static double g_errdom[33];
static double g_cur[33];
#define GYNC_ERRW(i) * g_errw[i]
#define GYNC_ERRSTAT(i, r) { g_cur[i] = (r) * (r); if (i == 32) { double t = 0; int a = 0; for (int k = 0; k < 33; ++k) { t += g_cur[k]; if (g_cur[k] > g_cur[a]) a = k; } \
                              if (t > 0) { for (int k = 0; k < 33; ++k) g_errshare[k] += g_cur[k] / t; g_errdom[a] += 1; } } }
#include "../../tests/host_shim/host_shim.cpp"
extern "C" {
void ew_set(const double* w) { memcpy(g_errw, w, sizeof(g_errw)); }
void ew_get(double* share, double* dom) { memcpy(share, g_errshare, sizeof(g_errshare)); memcpy(dom, g_errdom, sizeof(g_errdom)); }
void ew_reset() { memset(g_errshare, 0, sizeof(g_errshare)); memset(g_errdom, 0, sizeof(g_errdom)); }
}
