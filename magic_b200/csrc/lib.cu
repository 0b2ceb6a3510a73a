// lib.cu -- unity build of libmagic_b200.so (kernels are defined in headers; one translation unit keeps
// them in a single device link).  Build: see magic_b200/build.py.
#include "engine.cu"
#include "api_sht.cu"
#include "api_rloop.cu"
#include "api_diag.cu"
#include "api_transp.cu"
