/* Minimal stand-in for <Rinternals.h>: declarations only, so that reference headers which
 * mention SEXP types parse. None of these is ever called by the oracle harness; the handful that the
 * linker still asks for are defined as aborting stubs in oracle/ref_stubs.cpp. Test infrastructure only. */
#ifndef ORACLE_SHIM_RINTERNALS_H
#define ORACLE_SHIM_RINTERNALS_H
#include <R.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef int PROTECT_INDEX;
typedef unsigned int SEXPTYPE;
typedef struct R_inpstream_st *R_inpstream_t;
typedef struct R_outpstream_st *R_outpstream_t;
enum { NILSXP = 0, SYMSXP = 1, LISTSXP = 2, CLOSXP = 3, ENVSXP = 4, PROMSXP = 5, LANGSXP = 6, CHARSXP = 9, LGLSXP = 10,
       INTSXP = 13, REALSXP = 14, CPLXSXP = 15, STRSXP = 16, VECSXP = 19, EXPRSXP = 20, RAWSXP = 24 };
extern SEXP R_NilValue, R_UnboundValue, R_NamesSymbol, R_DimNamesSymbol, R_DimSymbol, R_ClassSymbol, R_RowNamesSymbol,
    R_LevelsSymbol, R_GlobalEnv, R_NaString, R_BlankString, R_MissingArg;
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
int Rf_length(SEXP);
R_xlen_t Rf_xlength(SEXP);
#define LENGTH(x) Rf_length(x)
#define XLENGTH(x) Rf_xlength(x)
int TYPEOF(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
double *REAL(SEXP);
int *INTEGER(SEXP);
int *LOGICAL(SEXP);
unsigned char *RAW(SEXP);
const char *CHAR(SEXP);
SEXP TAG(SEXP);
SEXP CAR(SEXP);
SEXP CDR(SEXP);
SEXP CADR(SEXP);
SEXP SETCAR(SEXP, SEXP);
void SET_TAG(SEXP, SEXP);
SEXP R_BytecodeExpr(SEXP);
Rboolean Rf_isReal(SEXP);
Rboolean Rf_isLogical(SEXP);
Rboolean Rf_isInteger(SEXP);
Rboolean Rf_isString(SEXP);
Rboolean Rf_isNewList(SEXP);
Rboolean Rf_isMatrix(SEXP);
Rboolean Rf_isNumeric(SEXP);
Rboolean Rf_isNull(SEXP);
Rboolean Rf_isFactor(SEXP);
Rboolean Rf_isEnvironment(SEXP);
Rboolean Rf_isVector(SEXP);
Rboolean Rf_isFunction(SEXP);
Rboolean Rf_isSymbol(SEXP);
Rboolean Rf_isLanguage(SEXP);
Rboolean Rf_inherits(SEXP, const char *);
SEXP Rf_install(const char *);
SEXP Rf_mkChar(const char *);
SEXP Rf_mkString(const char *);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_findVar(SEXP, SEXP);
SEXP Rf_findVarInFrame(SEXP, SEXP);
SEXP R_getVar(SEXP, SEXP, Rboolean);
SEXP R_getVarEx(SEXP, SEXP, Rboolean, SEXP);
void Rf_defineVar(SEXP, SEXP, SEXP);
void Rf_setVar(SEXP, SEXP, SEXP);
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocList(int);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_lcons(SEXP, SEXP);
SEXP Rf_cons(SEXP, SEXP);
SEXP Rf_lang1(SEXP);
SEXP Rf_lang2(SEXP, SEXP);
SEXP Rf_lang3(SEXP, SEXP, SEXP);
SEXP Rf_lang4(SEXP, SEXP, SEXP, SEXP);
SEXP Rf_eval(SEXP, SEXP);
SEXP R_tryEval(SEXP, SEXP, int *);
SEXP Rf_duplicate(SEXP);
SEXP Rf_coerceVector(SEXP, SEXPTYPE);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarLogical(int);
SEXP Rf_ScalarString(SEXP);
SEXP Rf_GetOption1(SEXP);
SEXP Rf_GetOption(SEXP, SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int Rf_asLogical(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
void R_PreserveObject(SEXP);
void R_ReleaseObject(SEXP);
void R_Serialize(SEXP, R_outpstream_t);
SEXP R_Unserialize(R_inpstream_t);
#ifdef __cplusplus
}
#endif
#endif
