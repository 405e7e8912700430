// general.h -- host-side configuration and geometry, same schema and semantics as the
// reference's src/general.{h,cpp} (flat input.json with 9 keys, .xyz geometry in Angstrom).
// Libint's Atom / BasisSet are replaced by the minimal Atom / BasisSet below (basis.h).
#pragma once
#include <string>
#include <vector>

#include "tensor.h"

struct Atom {
    int atomic_number;
    double x, y, z;// bohr
};

struct params {// general.h:29-39
    std::string inputfile;
    std::string type;
    std::string ref;
    std::string basis;
    int charge;
    int multiplicity;
    int maxiter;
    real_t scf_conv;
    real_t cc_conv;
};

// general.cpp:12-54: required filename,type,basis,maxiter,scf_conv; optional ref (-> "RHF"),
// charge (-> 0, parsed but unused), multiplicity (-> 1), cc_conv (-> scf_conv)
params read_config_json(const std::string &config_file);
// general.cpp:57-68: only .xyz is accepted (std::invalid_argument otherwise); Angstrom -> bohr
std::vector<Atom> read_geometry(const std::string &filename);
void print_geometry(const std::vector<Atom> &atoms);
