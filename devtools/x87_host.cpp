// x87_host.cpp -- development/test aid: csrc/x87_nrm2.h (the device's emulation of OpenBLAS's x87 dnrm2 kernel) compiled for
// the host, so that tests/test_x87_nrm2.py can compare it with the real kernel (scipy.linalg.blas.dnrm2) and with
// `long double` arithmetic in the GPU-less build container. Never loaded by the library.
//   g++ -O2 -std=c++17 -shared -fPIC -ffp-contract=off -o devtools/_x87_host.so devtools/x87_host.cpp
#include <cstddef>

#include "../neo_planner_b200/csrc/x87_nrm2.h"

extern "C" void sim_x87_nrm2(int count, int n, const double *v, double *fast, double *exact)
{
    for (int k = 0; k < count; k++) {
        fast[k] = neo::x87_nrm2(n, v + (size_t)k * n);           // what the kernel calls (double-double fast path + fallback)
        exact[k] = neo::x87_nrm2_exact(n, v + (size_t)k * n);    // the integer emulation alone
    }
}
