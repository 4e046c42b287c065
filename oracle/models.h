/* models.h -- ORACLE (test infrastructure, see drj_oracle.h): model registry of the CPU oracle.
 * Signature mirrors the reference's user-function contract, src/main.cpp:11-12
 * (params, data, misc) with misc["block"] (src/Particle.h:110) delivered as an int. */
#ifndef DRJ_ORACLE_MODELS_H
#define DRJ_ORACLE_MODELS_H
#include "drj_oracle.h"

typedef double (*ora_loglike_fn)(const double* theta, int d, const ora_list* data, const ora_list* misc, int block);
typedef double (*ora_logprior_fn)(const double* theta, int d, const ora_list* misc);

typedef struct ora_model {
  const char* name;
  ora_loglike_fn loglike;
  ora_logprior_fn logprior;
} ora_model;

const ora_model* ora_find_model(const char* name);
#endif
