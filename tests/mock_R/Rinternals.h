/*
 * Mock of the slice of R's C API that r/src/scdd_glue.c uses -- TEST INFRASTRUCTURE.
 *
 * R is not installed in the build image or on the GPU box, so the .Call glue could never meet a compiler.  This
 * header + mock_runtime.c let gcc compile it with -Wall -Wextra -Werror and let the tests EXECUTE it: vectors are
 * heap blocks, PROTECT/UNPROTECT keep a depth counter (balance is asserted), Rf_error longjmps back to the test
 * driver, R_registerRoutines records the table (arities are checked against the definitions).
 * Names, argument orders and semantics follow "Writing R Extensions" (sections 5.9-5.10, 6).
 */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mock_sexp* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
typedef enum { FALSE = 0, TRUE } Rboolean;

#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define RAWSXP 24

extern SEXP R_NilValue;
extern int R_NaInt;
#define NA_INTEGER R_NaInt
#define NA_LOGICAL R_NaInt
extern double R_NaReal;                  /* R's NA_real_: a NaN whose low word is 1954 */
#define NA_REAL R_NaReal
#define ISNAN(x) (isnan(x) != 0)         /* R_ext/Arith.h */

SEXP Rf_allocVector(unsigned int type, R_xlen_t n);
SEXP Rf_allocMatrix(unsigned int type, int nrow, int ncol);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
double* REAL(SEXP);
int* INTEGER(SEXP);
int* LOGICAL(SEXP);
Rbyte* RAW(SEXP);
int LENGTH(SEXP);
R_xlen_t XLENGTH(SEXP);
int TYPEOF(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
Rboolean Rf_isNull(SEXP);
SEXP Rf_mkNamed(unsigned int type, const char** names);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarLogical(int);
SEXP Rf_ScalarReal(double);
char* R_alloc(size_t n, int size);
void Rf_error(const char* fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
void Rf_warning(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
void R_CheckUserInterrupt(void);
Rboolean R_ToplevelExec(void (*fun)(void*), void* data);

#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
#define asInteger Rf_asInteger
#define asLogical Rf_asLogical
#define asReal Rf_asReal
#define isNull Rf_isNull
#define mkNamed Rf_mkNamed
#define ScalarInteger Rf_ScalarInteger
#define ScalarLogical Rf_ScalarLogical
#define ScalarReal Rf_ScalarReal
#define error Rf_error
#define warning Rf_warning

#ifdef __cplusplus
}
#endif
#endif
