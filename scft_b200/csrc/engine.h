// engine.h — internal declarations shared by the C-ABI translation units.
#pragma once
#include <string>
#include <vector>

#include "../../include/scft_b200.h"

namespace scftb {
int fail(int code, const std::string &msg);
int romberg_weights(int m, double hh, std::vector<double> &w);
void trapezoid_weights(int m, double hh, std::vector<double> &w);
void f0_given(int N, const double *x, double tau, double *f0);
int gauss_jordan(std::vector<double> &a, int n, std::vector<double> &b, bool nudge);
}  // namespace scftb

// internal accessor (not part of the public ABI)
extern "C" int scftb_engine_max_batch(scftb_engine *e);
