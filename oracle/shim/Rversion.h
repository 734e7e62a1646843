#ifndef ORACLE_SHIM_RVERSION_H
#define ORACLE_SHIM_RVERSION_H
#define R_Version(v,p,s) (((v) * 65536) + ((p) * 256) + (s))
#define R_VERSION R_Version(4,5,0)
#endif
