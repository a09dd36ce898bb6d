/* oracle/gsl_shim.h -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference (pelahi/NBodylib) needs GSL headers to *parse* (NBodyMath.h drags in
 * fitting/integration headers) but the kd-tree path calls exactly one GSL symbol,
 * gsl_sf_gamma (reference src/KDTree/KDTree.cxx:1158).  GSL is not installed in this image,
 * so the recipe in oracle/Makefile points -I at a directory of one-line headers that all
 * include this file.  std::tgamma(ND/2+1) reproduces kernnorm to the last bit for ND=3,6.
 */
#ifndef NBK_ORACLE_GSL_SHIM_H
#define NBK_ORACLE_GSL_SHIM_H
#include <cmath>
#include <cstddef>
#define GSL_SUCCESS 0
struct gsl_vector; struct gsl_matrix; struct gsl_monte_function; struct gsl_function; struct gsl_rng;
struct gsl_multifit_nlinear_fdf; struct gsl_multifit_nlinear_workspace; struct gsl_multifit_nlinear_parameters;
static inline const char* gsl_strerror(int) { return "gsl shim"; }
static inline double gsl_sf_gamma(double x) { return std::tgamma(x); }
#endif
