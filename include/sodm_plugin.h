/* sodm_plugin.h -- C ABI of a MODEL PLUGIN: one user model translated to a device functor and compiled, by nvcc for
 * sm_100a, into its own shared library next to libsodm.so.
 *
 * The reference runs arbitrary Python model functions (`integrate`, `compute_rates`, `apply_transitionings`,
 * /root/reference/src/pySODM/models/base.py:352-417, :746-796).  `pysodm_b200.codegen` traces such a function into
 * straight-line C++, wraps it in a functor with the interface of the built-in ones and instantiates the SAME kernels
 * (csrc/jump_kernels.cuh, csrc/ode_kernels.cuh).  A plugin exports exactly these symbols; the argument structs and all
 * conventions (caller-owned device buffers, asynchronous on `stream`, negative return = SODM_E* with a message from
 * sodm_plugin_last_error) are those of sodm.h.  `args->model` is ignored: a plugin holds one model.
 */
#ifndef SODM_PLUGIN_H
#define SODM_PLUGIN_H

#include "sodm.h"

#ifdef __cplusplus
extern "C" {
#endif

int         sodm_plugin_abi_version(void);                 /* == SODM_ABI_VERSION of the headers it was built with   */
const char* sodm_plugin_last_error(void);
int         sodm_plugin_info(sodm_model_info* out);        /* kind, flat sizes, channels, time-dependent range        */
/* Exactly one of the two does work, according to sodm_plugin_info().kind; the other returns SODM_EUNSUPPORTED. */
int         sodm_plugin_jump_run(const sodm_jump_args* args);   /* replaces sodm_jump_run for this model             */
int         sodm_plugin_ode_run(const sodm_ode_args* args);     /* replaces sodm_ode_run for this model              */

#ifdef __cplusplus
}
#endif
#endif /* SODM_PLUGIN_H */
