// backend.h -- TEST INFRASTRUCTURE ONLY (oracle/ref_shim): flat C interface through which the
// libint2 shim reaches the repo's host integral substrate (qc-p2_b200/host/{basis,gaussian}.cpp).
// Built as its own shared object with hidden visibility + -Bsymbolic, because the host substrate
// and the reference define global functions of the same names (read_config_json, density_guess, ...).
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
void *shimb_basis_create(const char *name, int natom, const int *Z, const double *xyz_bohr);// NULL on error
void shimb_basis_destroy(void *h);
int shimb_nshell(void *h);
int shimb_nbf(void *h);
// l, nprim, origin[3]; alpha/coef receive up to maxprim entries.  Returns nprim.
int shimb_shell(void *h, int s, int *l, double *origin, double *alpha, double *coef, int maxprim);
// op: 0 overlap, 1 kinetic, 2 nuclear (point charges = the atoms given); out[nbf*nbf] row-major
int shimb_onebody(void *h, int op, int natom, const int *Z, const double *xyz_bohr, double *out);
int shimb_eri(void *h, double *out);// out[nbf^4], (pq|rs)
#ifdef __cplusplus
}
#endif
