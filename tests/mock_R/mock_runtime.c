/* Mock R runtime behind tests/mock_R/Rinternals.h -- TEST INFRASTRUCTURE (see that header).
 * Also exports the small driver API the Python tests use to build arguments and read results (mock_*). */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R.h"
#include "R_ext/Rdynload.h"

struct mock_sexp {
    int type;
    R_xlen_t n;
    void* data;
    int nrow, ncol;          /* matrices */
    char** names;            /* VECSXP made by mkNamed */
};

static struct mock_sexp nil_obj = {NILSXP, 0, NULL, 0, 0, NULL};
SEXP R_NilValue = &nil_obj;
int R_NaInt = (-2147483647 - 1);
double R_NaReal;

__attribute__((constructor)) static void mock_init_na(void) {
    union { double d; unsigned long long u; } x;
    x.u = 0x7FF00000000007A2ULL;         /* arithmetic.c R_ValueOfNA: hw = 0x7FF00000, lw = 1954 */
    R_NaReal = x.d;
}

static int protect_depth = 0, protect_max = 0, protect_underflow = 0;
static long n_alloc = 0;
static char err_msg[1024];
static jmp_buf* err_jmp = NULL;
static int interrupt_pending = 0, interrupt_after_checks = -1, n_interrupt_checks = 0;
static const R_CallMethodDef* reg_table = NULL;

static size_t elt_size(unsigned int type) {
    switch (type) {
        case LGLSXP: case INTSXP: return sizeof(int);
        case REALSXP: return sizeof(double);
        case RAWSXP: return 1;
        case VECSXP: case STRSXP: return sizeof(SEXP);
        default: return 0;
    }
}

SEXP Rf_allocVector(unsigned int type, R_xlen_t n) {
    SEXP s = (SEXP)calloc(1, sizeof(struct mock_sexp));
    s->type = (int)type; s->n = n;
    s->data = calloc((size_t)(n > 0 ? n : 1), elt_size(type) ? elt_size(type) : 1);
    n_alloc++;
    return s;
}
SEXP Rf_allocMatrix(unsigned int type, int nrow, int ncol) {
    SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    s->nrow = nrow; s->ncol = ncol;
    return s;
}
SEXP Rf_protect(SEXP s) { protect_depth++; if (protect_depth > protect_max) protect_max = protect_depth; return s; }
void Rf_unprotect(int n) { protect_depth -= n; if (protect_depth < 0) { protect_underflow = 1; protect_depth = 0; } }
static void* typed(SEXP s, int t1, int t2, const char* what) {
    if (s->type != t1 && s->type != t2) Rf_error("%s() applied to an object of type %d", what, s->type);
    return s->data;
}
double* REAL(SEXP s) { return (double*)typed(s, REALSXP, REALSXP, "REAL"); }
int* INTEGER(SEXP s) { return (int*)typed(s, INTSXP, LGLSXP, "INTEGER"); }   /* as R: INTEGER() on a double vector is an error */
int* LOGICAL(SEXP s) { return (int*)typed(s, LGLSXP, LGLSXP, "LOGICAL"); }
Rbyte* RAW(SEXP s) { return (Rbyte*)typed(s, RAWSXP, RAWSXP, "RAW"); }
int LENGTH(SEXP s) { return (int)s->n; }
R_xlen_t XLENGTH(SEXP s) { return s->n; }
int TYPEOF(SEXP s) { return s->type; }
int Rf_asInteger(SEXP s) {
    if (s->n < 1) return R_NaInt;
    if (s->type == INTSXP || s->type == LGLSXP) return ((int*)s->data)[0];
    if (s->type == REALSXP) return (int)((double*)s->data)[0];
    return R_NaInt;
}
int Rf_asLogical(SEXP s) { int v = Rf_asInteger(s); return v == R_NaInt ? v : (v != 0); }
double Rf_asReal(SEXP s) {
    if (s->n < 1) return 0.0 / 0.0;
    if (s->type == REALSXP) return ((double*)s->data)[0];
    return (double)Rf_asInteger(s);
}
Rboolean Rf_isNull(SEXP s) { return s == R_NilValue || s->type == NILSXP; }
SEXP Rf_mkNamed(unsigned int type, const char** names) {
    int n = 0;
    while (names[n][0] != '\0') n++;
    SEXP s = Rf_allocVector(type, n);
    s->names = (char**)calloc((size_t)n + 1, sizeof(char*));
    for (int i = 0; i < n; i++) s->names[i] = strdup(names[i]);
    return s;
}
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) {
    if (x->type != VECSXP || i < 0 || i >= x->n) Rf_error("SET_VECTOR_ELT out of range");
    ((SEXP*)x->data)[i] = v;
    return v;
}
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) {
    if (x->type != VECSXP || i < 0 || i >= x->n) Rf_error("VECTOR_ELT out of range");
    return ((SEXP*)x->data)[i];
}
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); ((int*)s->data)[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); ((int*)s->data)[0] = v; return s; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); ((double*)s->data)[0] = v; return s; }
char* R_alloc(size_t n, int size) { return (char*)calloc(n ? n : 1, (size_t)(size > 0 ? size : 1)); }   /* freed with the process */
void Rf_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_msg, sizeof(err_msg), fmt, ap);
    va_end(ap);
    if (err_jmp) longjmp(*err_jmp, 1);
    fprintf(stderr, "mock R error outside mock_call: %s\n", err_msg);
    abort();
}
void Rf_warning(const char* fmt, ...) { (void)fmt; }
void R_CheckUserInterrupt(void) {
    n_interrupt_checks++;
    if (interrupt_after_checks >= 0 && n_interrupt_checks > interrupt_after_checks) interrupt_pending = 1;
    if (interrupt_pending) Rf_error("interrupted");      /* R longjmps to top level */
}
Rboolean R_ToplevelExec(void (*fun)(void*), void* data) {
    /* runs fun; FALSE if it jumped (error / interrupt) */
    jmp_buf here, *saved = err_jmp;
    err_jmp = &here;
    Rboolean ok = TRUE;
    if (setjmp(here) == 0) fun(data); else ok = FALSE;
    err_jmp = saved;
    return ok;
}
int R_registerRoutines(DllInfo* info, const R_CMethodDef* const c, const R_CallMethodDef* const call, const void* const f,
                       const void* const e) {
    (void)info; (void)c; (void)f; (void)e;
    reg_table = call;
    return 1;
}
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value) { (void)info; (void)value; return TRUE; }
Rboolean R_forceSymbols(DllInfo* info, Rboolean value) { (void)info; (void)value; return TRUE; }

/* ------------------------------------------------------------------ driver API for the tests (ctypes) */
void R_init_scDD(DllInfo* dll);

SEXP mock_real(const double* v, long n) { SEXP s = Rf_allocVector(REALSXP, n); if (n) memcpy(s->data, v, (size_t)n * 8); return s; }
SEXP mock_int(const int* v, long n) { SEXP s = Rf_allocVector(INTSXP, n); if (n) memcpy(s->data, v, (size_t)n * 4); return s; }
SEXP mock_real_matrix(const double* v, int nrow, int ncol) { SEXP s = mock_real(v, (long)nrow * ncol); s->nrow = nrow; s->ncol = ncol; return s; }
SEXP mock_int_matrix(const int* v, int nrow, int ncol) { SEXP s = mock_int(v, (long)nrow * ncol); s->nrow = nrow; s->ncol = ncol; return s; }
SEXP mock_lgl(int v) { return Rf_ScalarLogical(v); }
SEXP mock_nil(void) { return R_NilValue; }
int mock_type(SEXP s) { return s->type; }
long mock_len(SEXP s) { return (long)s->n; }
int mock_nrow(SEXP s) { return s->nrow; }
int mock_ncol(SEXP s) { return s->ncol; }
void* mock_data(SEXP s) { return s->data; }
SEXP mock_list_get(SEXP s, const char* name) {
    if (s->type != VECSXP || !s->names) return NULL;
    for (R_xlen_t i = 0; i < s->n; i++) if (strcmp(s->names[i], name) == 0) return ((SEXP*)s->data)[i];
    return NULL;
}
int mock_protect_depth(void) { return protect_depth; }
int mock_protect_underflow(void) { return protect_underflow; }
const char* mock_last_error(void) { return err_msg; }
void mock_interrupt_after(int checks) { interrupt_after_checks = checks; n_interrupt_checks = 0; interrupt_pending = 0; }
int mock_interrupt_checks(void) { return n_interrupt_checks; }
int mock_registered(const char* name) {   /* numArgs of a registered .Call routine, -1 if absent */
    if (!reg_table) R_init_scDD(NULL);
    for (const R_CallMethodDef* r = reg_table; r && r->name; r++) if (strcmp(r->name, name) == 0) return r->numArgs;
    return -1;
}
DL_FUNC mock_routine(const char* name) {
    if (!reg_table) R_init_scDD(NULL);
    for (const R_CallMethodDef* r = reg_table; r && r->name; r++) if (strcmp(r->name, name) == 0) return r->fun;
    return NULL;
}
/* .Call(name, args...): looks the routine up in the registered table, checks the arity, traps Rf_error.
 * Returns NULL on error (message in mock_last_error()). */
typedef SEXP (*call0)(void);
typedef SEXP (*call10)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP mock_call(const char* name, int nargs, SEXP* a) {
    DL_FUNC f = mock_routine(name);
    err_msg[0] = '\0';
    if (!f) { snprintf(err_msg, sizeof(err_msg), "no such routine %s", name); return NULL; }
    if (mock_registered(name) != nargs) { snprintf(err_msg, sizeof(err_msg), "%s registered with %d args, called with %d", name, mock_registered(name), nargs); return NULL; }
    if (nargs > 10) { snprintf(err_msg, sizeof(err_msg), "too many args"); return NULL; }
    SEXP v[10];
    for (int i = 0; i < 10; i++) v[i] = i < nargs ? a[i] : R_NilValue;
    jmp_buf here;
    err_jmp = &here;
    SEXP out = NULL;
    const int depth0 = protect_depth;
    if (setjmp(here) == 0) {
        /* the x86-64 / aarch64 C calling conventions allow calling a function of fewer SEXP parameters with more */
        out = ((call10)f)(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9]);
    } else {
        protect_depth = depth0;       /* R unwinds the protect stack on error */
        out = NULL;
    }
    err_jmp = NULL;
    return out;
}
