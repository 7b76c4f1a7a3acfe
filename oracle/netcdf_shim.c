/* TEST INFRASTRUCTURE (oracle/): link shim for the 15 NetCDF symbols the reference's
 * Common/IO.h and Grid3D.cpp reference (src/Common/IO.h:136-388, Grid3D.cpp:420-486).
 * libnetcdf is not installed in this image; the probe drivers never write NetCDF, so
 * every entry point is a no-op that reports success (nc_open fails: SeaNetCDF inputs
 * are unreadable here and excluded from parity). */
#include <stddef.h>
int nc_create(const char *p, int m, int *id) { (void)p; (void)m; if (id) *id = 1; return 0; }
int nc_open(const char *p, int m, int *id) { (void)p; (void)m; (void)id; return -1; }
int nc_close(int id) { (void)id; return 0; }
int nc_def_dim(int id, const char *n, size_t l, int *d) { (void)id; (void)n; (void)l; if (d) *d = 0; return 0; }
int nc_def_var(int id, const char *n, int t, int nd, const int *ds, int *v) { (void)id; (void)n; (void)t; (void)nd; (void)ds; if (v) *v = 0; return 0; }
int nc_enddef(int id) { (void)id; return 0; }
int nc_get_var(int id, int v, void *ip) { (void)id; (void)v; (void)ip; return -1; }
int nc_inq_dimid(int id, const char *n, int *d) { (void)id; (void)n; (void)d; return -1; }
int nc_inq_dimlen(int id, int d, size_t *l) { (void)id; (void)d; (void)l; return -1; }
int nc_inq_varid(int id, const char *n, int *v) { (void)id; (void)n; if (v) *v = 0; return 0; }
int nc_put_att_double(int id, int v, const char *n, int t, size_t l, const double *op) { (void)id; (void)v; (void)n; (void)t; (void)l; (void)op; return 0; }
int nc_put_att_float(int id, int v, const char *n, int t, size_t l, const float *op) { (void)id; (void)v; (void)n; (void)t; (void)l; (void)op; return 0; }
int nc_put_att_text(int id, int v, const char *n, size_t l, const char *op) { (void)id; (void)v; (void)n; (void)l; (void)op; return 0; }
int nc_put_var_float(int id, int v, const float *op) { (void)id; (void)v; (void)op; return 0; }
int nc_put_vara_double(int id, int v, const size_t *s, const size_t *c, const double *op) { (void)id; (void)v; (void)s; (void)c; (void)op; return 0; }
