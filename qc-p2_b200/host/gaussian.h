// gaussian.h -- host integrals (stand-in for the Libint engines used in hf.cpp:37-88 and
// integrals.cpp:12-89); see gaussian.cpp.
#pragma once
#include <vector>

#include "basis.h"
#include "tensor.h"

enum class Operator1 { overlap, kinetic, nuclear };

double compute_enuc(const std::vector<Atom> &atoms);                         // hf.cpp:21-34
Matrix compute_1body_ints(const BasisSet &obs, Operator1 op,                  // hf.cpp:37-88
                          const std::vector<Atom> &atoms = std::vector<Atom>());
DTensor eri_ao_tensor(const BasisSet &obs);                                   // integrals.cpp:12-89
double contracted_norm(const Shell &s);                                       // self-test of the built-in basis data
