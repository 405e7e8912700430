// mp2.cpp -- mp2.h entry points on the GPU (reference bodies: mp2.cpp:11-86).
#include "mp2.h"

#include <iostream>

#include "gpu.h"

DTensor make_so_moes(const Vector &eps_a, const Vector &eps_b, const int &nao) {
    DTensor eps_so(2 * nao, 2 * nao);
    gpu_check(qcp2_make_so_moes(gpu(), eps_a.data(), eps_b.data(), nao, eps_so.data()));
    return eps_so;
}

mp2_output run_mp2(const DTensor &oovv, const DTensor &denom) {
    mp2_output out;
    out.T = DTensor(oovv.extent(0), oovv.extent(1), oovv.extent(2), oovv.extent(3));
    gpu_check(qcp2_run_mp2(gpu(), oovv.data(), denom.data(), (int) oovv.extent(0), (int) oovv.extent(2), out.T.data(), &out.energy));
    return out;
}

mp2_results MP2(const BasisSet &obs, const scf_results &scf, const params &config, const DTensor *ao_in) {
    mp2_results results;
    std::cout << std::endl << "Starting MP2 calculation" << std::endl << std::endl;
    DTensor own;
    if (!ao_in) own = eri_ao_tensor(obs), ao_in = &own;
    const bool rhf = config.ref == "RHF" || config.ref == "rhf";
    const bool uhf = config.ref == "UHF" || config.ref == "uhf";
    if (!rhf && !uhf) throw std::invalid_argument("MP2: unknown reference " + config.ref);
    const int n = scf.nao;
    results.so_ints = DTensor(2 * n, 2 * n, 2 * n, 2 * n);
    results.T = DTensor(scf.no, scf.no, scf.nv, scf.nv);
    // one fused device pipeline: AO->MO (x1 RHF, x3 UHF as written), spin-orbital gather, oovv
    // slice, denominators, amplitudes and energy (mp2.cpp:50-83)
    gpu_check(qcp2_mp2(gpu(), ao_in->data(), rhf ? scf.C.data() : scf.Ca.data(), rhf ? nullptr : scf.Cb.data(),
                       rhf ? scf.moes.data() : scf.moes_a.data(), rhf ? nullptr : scf.moes_b.data(), n, scf.no, scf.nv,
                       uhf ? 1 : 0, results.so_ints.data(), results.T.data(), &results.energy));
    std::cout << "MP2 Energy: " << results.energy << " Eh" << std::endl;
    return results;
}
