/* Minimal stand-in for <R.h>: lets the reference's hot-path sources compile where R is absent.
 * Test infrastructure only (see oracle/README.md). Declarations, no behaviour. */
#ifndef ORACLE_SHIM_R_H
#define ORACLE_SHIM_R_H
#include <stddef.h>
#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <math.h>
#include <limits.h>
#include <R_ext/Arith.h>
#ifdef __cplusplus
extern "C" {
#endif
void Rprintf(const char *, ...);
void REprintf(const char *, ...);
void Rf_error(const char *, ...);
void Rf_warning(const char *, ...);
void R_CheckUserInterrupt(void);
double unif_rand(void);
void GetRNGstate(void);
void PutRNGstate(void);
void R_FlushConsole(void);
#ifdef __cplusplus
}
#endif
#endif
