// oracle/ref_stubs.cpp -- TEST INFRASTRUCTURE ONLY.
// Definitions for the few R-runtime / rdb:: symbols the reference's hot-path objects reference at link time.
// None is reachable from the paths ref_harness.cpp drives except rdb::verror (turned into an exception)
// and rdb::get_chrom_files (a readdir).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <dirent.h>
#include <stdexcept>
#include <string>
#include <vector>
#include <Rinternals.h>

namespace rdb {
void verror(const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	throw std::runtime_error(buf);
}
void rerror(const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	throw std::runtime_error(buf);
}
void check_interrupt() {}
void get_chrom_files(const char *dirname, std::vector<std::string> &chrom_files)
{
	chrom_files.clear();
	DIR *d = opendir(dirname);
	if (!d)
		return;
	while (struct dirent *e = readdir(d))
		if (e->d_name[0] != '.')
			chrom_files.push_back(e->d_name);
	closedir(d);
}
} // namespace rdb

static void unreachable(const char *what)
{
	fprintf(stderr, "oracle/ref_stubs: R runtime symbol %s reached - the harness must never get here\n", what);
	abort();
}

extern "C" {
SEXP R_NilValue = 0, R_DimNamesSymbol = 0, R_NamesSymbol = 0, R_UnboundValue = 0;
int Rf_nrows(SEXP) { unreachable("Rf_nrows"); return 0; }
SEXP Rf_getAttrib(SEXP, SEXP) { unreachable("Rf_getAttrib"); return 0; }
Rboolean Rf_isNull(SEXP) { unreachable("Rf_isNull"); return FALSE; }
SEXP VECTOR_ELT(SEXP, R_xlen_t) { unreachable("VECTOR_ELT"); return 0; }
Rboolean Rf_isString(SEXP) { unreachable("Rf_isString"); return FALSE; }
int Rf_length(SEXP) { unreachable("Rf_length"); return 0; }
const char *CHAR(SEXP) { unreachable("CHAR"); return 0; }
SEXP STRING_ELT(SEXP, R_xlen_t) { unreachable("STRING_ELT"); return 0; }
double *REAL(SEXP) { unreachable("REAL"); return 0; }
double unif_rand(void) { unreachable("unif_rand"); return 0; }
void Rf_warning(const char *, ...) {}
void Rf_error(const char *fmt, ...) { unreachable("Rf_error"); }
void Rprintf(const char *, ...) {}
void REprintf(const char *, ...) {}
void R_CheckUserInterrupt(void) {}
}
