// main.cpp -- `P2 <input.json>`: same dispatch on config.type / config.ref as the reference's
// main.cpp:13-95; SCF on the host, MP2 / CCSD / (T) on the GPU through the C ABI.  Besides the
// reference's setprecision(12) prints it writes machine-readable results with %.17g
// (12 digits on a -76 Eh total are too coarse to assert 1e-10).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
    // `P2 --fused <input.json>`: same inputs and outputs, but MP2 -> CCSD -> (T) run as ONE device-
    // resident call (qcp2_post_hf) instead of the reference's function-by-function sequence, which
    // moves so_ints / sliced integrals through host memory between the stages.
    // `P2 --gpus N <input.json>`: N processes, one per GPU (forked here, before any CUDA / OpenMP
    // state exists); rank 0 runs the program below, the others join the (T) stage over NCCL.
    bool fused = false;
    int gpus = 1;
    for (;;) {
        if (argc >= 2 && std::string(argv[1]) == "--fused") fused = true, --argc, ++argv;
        else if (argc >= 3 && std::string(argv[1]) == "--gpus") gpus = std::atoi(argv[2]), argc -= 2, argv += 2;
        else break;
    }
    if (argc < 2 || gpus < 1) {
        std::cerr << "usage: P2 [--fused] [--gpus N] <input.json>" << endl;
        return 2;
    }
    try {
        mgpu_spawn(gpus);
    } catch (const std::exception &e) {
        std::cerr << "P2: " << e.what() << endl;
        return 1;
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
            // the reference matches the literal strings only (main.cpp:58,70,80; cc.cpp:39,52): anything else silently runs no
            // SCF there and continues on uninitialised results -- here it is refused with the reason
            if (!is_rhf && !is_uhf)
                throw std::runtime_error("unsupported \"ref\": \"" + config.ref + "\" (use \"RHF\" or \"UHF\", upper case as the reference requires)");
            if (is_rhf) hf = RHF(atoms, obs, nelectron, config, &ao);
            else if (is_uhf) hf = UHF(atoms, obs, nelectron, config, &ao);
            if (!hf.converged) throw std::runtime_error("SCF did not converge within maxiter (the reference leaves its results undefined here)");
            e_scf = hf.energy;
            if (fused) {
                const int stages = (t == "MP2" || t == "mp2") ? 1 : ((t == "CCSD" || t == "ccsd") ? 2 : 3);
                int iters = 0;
                std::vector<double> hist((size_t) std::max(config.maxiter, 1), 0.0);
                if (mgpu().world > 1) {// every rank joins the collective pipeline: sharded CCSD products, round-robin (T)
                    MgpuCommand cmd;
                    std::memset(&cmd, 0, sizeof(cmd));
                    cmd.cmd = 2, cmd.o = hf.no, cmd.v = hf.nv, cmd.n = hf.nao, cmd.uhf = is_uhf ? 1 : 0, cmd.stages = stages;
                    cmd.maxiter = config.maxiter, cmd.conv = config.cc_conv;
                    if (qcp2_comm_get_unique_id(cmd.id, QCP2_COMM_ID_BYTES) != 0) throw std::runtime_error("qcp2_b200: NCCL is not available");
                    mgpu_send(cmd);
                    gpu_check(qcp2_comm_init(gpu(), cmd.id, mgpu().world, 0));
                    gpu_check(qcp2_post_hf_mgpu(gpu(), ao.data(), is_rhf ? hf.C.data() : hf.Ca.data(), is_rhf ? nullptr : hf.Cb.data(),
                                                is_rhf ? hf.moes.data() : hf.moes_a.data(), is_rhf ? nullptr : hf.moes_b.data(), hf.nao,
                                                hf.no, hf.nv, is_uhf ? 1 : 0, stages, config.maxiter, config.cc_conv, 0, &e_mp2, &e_cc, &e_t,
                                                &iters, hist.data(), nullptr, nullptr));
                    gpu_check(qcp2_comm_destroy(gpu()));
                } else
                gpu_check(qcp2_post_hf(gpu(), ao.data(), is_rhf ? hf.C.data() : hf.Ca.data(), is_rhf ? nullptr : hf.Cb.data(),
                                       is_rhf ? hf.moes.data() : hf.moes_a.data(), is_rhf ? nullptr : hf.moes_b.data(), hf.nao, hf.no,
                                       hf.nv, is_uhf ? 1 : 0, stages, config.maxiter, config.cc_conv, &e_mp2, &e_cc, &e_t, &iters,
                                       hist.data(), nullptr, nullptr));
                cout << "MP2 Energy: " << e_mp2 << " Eh" << endl;
                if (stages >= 2) {
                    for (int it = 0; it <= std::min(iters, config.maxiter - 1); ++it) printf(" %02d %20.12f\n", it, hist[(size_t) it]);
                    cout << "CCSD energy: " << e_cc << " Eh" << endl;
                }
                if (stages >= 3) cout << "(T) energy: " << e_t << " Eh" << endl;
                cout << "Total energy: " << e_scf + (stages >= 2 ? e_cc : e_mp2) + e_t << " Eh" << endl;
                printf("RESULT {\"e_scf\": %.17g, \"e_mp2\": %.17g, \"e_ccsd\": %.17g, \"e_t\": %.17g}\n", e_scf, e_mp2, e_cc, e_t);
                gpu_release();
                return mgpu_shutdown();
            }
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
        mgpu_shutdown();
        return 1;
    }
    gpu_release();
    return mgpu_shutdown();
}
