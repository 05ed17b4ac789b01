/* Minimal stand-in for R's C API headers (R is not installed in this image), declaring ONLY what r-package/src/rshim.c
 * uses, with the signatures documented in "Writing R Extensions".  tests/test_abi_and_host.py type-checks the shim against
 * these with `gcc -fsyntax-only`: a typo, a wrong argument count against include/scmerge_b200.h or a missing symbol in the
 * shim fails on CPU even though the shim cannot be linked or run here. */
#ifndef R_STUB_R_H
#define R_STUB_R_H
#include <stddef.h>
#include <stdint.h>
char* R_alloc(size_t n, int size);
void* R_chk_calloc(size_t n, size_t size);
void R_chk_free(void* p);
#define R_Calloc(n, t) ((t*)R_chk_calloc((size_t)(n), sizeof(t)))
#define R_Free(p) (R_chk_free((void*)(p)), (p) = NULL)
#endif
