// mp2.h -- same declarations as the reference's src/mp2.h:12-27.
#pragma once
#include "hf.h"
#include "integrals.h"

struct mp2_results {
    real_t energy;
    DTensor so_ints;
    DTensor T;
};
struct mp2_output {
    real_t energy;
    DTensor T;
};

DTensor make_so_moes(const Vector &eps_a, const Vector &eps_b, const int &nao);
mp2_output run_mp2(const DTensor &oovv, const DTensor &denom);
// ao_ints: the AO tensor the SCF already built (the reference recomputes it, mp2.cpp:52/69)
mp2_results MP2(const BasisSet &obs, const scf_results &scf, const params &config, const DTensor *ao_ints = nullptr);
