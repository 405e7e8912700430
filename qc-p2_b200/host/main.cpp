// main.cpp -- `P2 <input.json>`: same dispatch on config.type / config.ref as the reference's
// main.cpp:13-95; SCF on the host, MP2 / CCSD / (T) on the GPU through the C ABI.  Besides the
// reference's setprecision(12) prints it writes machine-readable results with %.17g
// (12 digits on a -76 Eh total are too coarse to assert 1e-10).
#include <cstdio>
#include <iomanip>
#include <iostream>

#include "cc.h"
#include "general.h"
#include "gpu.h"
#include "hf.h"
#include "mp2.h"

int main(int argc, char *argv[]) {
    using std::cout;
    using std::endl;
    if (argc < 2) {
        std::cerr << "usage: P2 <input.json>" << endl;
        return 2;
    }
    cout << std::setprecision(12);
    try {
        auto config = read_config_json(argv[1]);
        std::vector<Atom> atoms = read_geometry(config.inputfile);
        auto nelectron = 0;
        for (auto &atom : atoms) nelectron += atom.atomic_number;// config.charge is parsed but never used (main.cpp:29-31)
        print_geometry(atoms);
        cout << "Number of electrons = " << nelectron << endl;
        BasisSet obs(config.basis, atoms);
        auto nao = nbasis(obs.shells());
        cout << endl << "Method: " << config.type << endl << "Reference: " << config.ref << endl
             << "Basis Set: " << config.basis << endl << "Number of basis functions = " << nao << endl;
        const std::string t = config.type;
        const bool is_rhf = config.ref == "RHF" || config.ref == "rhf", is_uhf = config.ref == "UHF";
        double e_scf = 0, e_mp2 = 0, e_cc = 0, e_t = 0;
        if (t == "RHF" || t == "rhf") {
            e_scf = RHF(atoms, obs, nelectron, config).energy;
        } else if (t == "UHF" || t == "uhf") {
            e_scf = UHF(atoms, obs, nelectron, config).energy;
        } else if (t == "MP2" || t == "mp2" || t == "CCSD" || t == "ccsd" || t == "CCSD(T)" || t == "ccsd(t)") {
            DTensor ao = eri_ao_tensor(obs);
            scf_results hf;
            if (is_rhf) hf = RHF(atoms, obs, nelectron, config, &ao);
            else if (is_uhf) hf = UHF(atoms, obs, nelectron, config, &ao);
            if (!hf.converged) throw std::runtime_error("SCF did not converge within maxiter (the reference leaves its results undefined here)");
            e_scf = hf.energy;
            auto mp2 = MP2(obs, hf, config, &ao);
            e_mp2 = mp2.energy;
            if (t == "MP2" || t == "mp2") {
                cout << "Total MP2 energy: " << e_scf + e_mp2 << " Eh" << endl;
            } else {
                auto cc = CCSD(hf, mp2, config);
                e_cc = cc.ccsd_energy;
                if (t == "CCSD(T)" || t == "ccsd(t)") {
                    auto m = make_moe_tensors(hf, config);
                    e_t = CCSD_T(cc, m);
                }
                cout << "Total energy: " << e_scf + e_cc + e_t << " Eh" << endl;
            }
        } else {
            cout << endl << config.type << " is not a supported method" << endl;
        }
        printf("RESULT {\"e_scf\": %.17g, \"e_mp2\": %.17g, \"e_ccsd\": %.17g, \"e_t\": %.17g}\n", e_scf, e_mp2, e_cc, e_t);
    } catch (const std::exception &e) {
        std::cerr << "P2: " << e.what() << endl;
        gpu_release();
        return 1;
    }
    gpu_release();
    return 0;
}
