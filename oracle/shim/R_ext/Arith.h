#ifndef ORACLE_SHIM_R_ARITH_H
#define ORACLE_SHIM_R_ARITH_H
#ifdef __cplusplus
#include <limits>
#include <cmath>
#define R_NegInf (-std::numeric_limits<double>::infinity())
#define R_PosInf (std::numeric_limits<double>::infinity())
#define R_NaN (std::numeric_limits<double>::quiet_NaN())
#define R_NaReal (std::numeric_limits<double>::quiet_NaN())
#define NA_REAL R_NaReal
#define NA_INTEGER INT_MIN
#define NA_LOGICAL INT_MIN
#define R_FINITE(x) std::isfinite(x)
#define ISNAN(x) std::isnan(x)
#endif
#endif
