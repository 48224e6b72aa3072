// Back-ends not built yet report PDEB200_EUNSUPPORTED (never a CPU fallback).
#include "ctx.hpp"
namespace pdeb200 {
#ifndef PDEB_HAVE_KSEG
int32_t kseg_setup(pdeb200_ctx* c) { return fail(c, PDEB200_EUNSUPPORTED, "Keller-Segel back-end not built"); }
int32_t kseg_core(pdeb200_ctx* c) { return fail(c, PDEB200_EUNSUPPORTED, "Keller-Segel back-end not built"); }
int32_t kseg_cost(const pdeb200_ctx*, double*, double*) { return PDEB200_EUNSUPPORTED; }
void kseg_free(pdeb200_ctx*) {}
#endif
#ifndef PDEB_HAVE_KSEG2D
int32_t kseg2d_setup(pdeb200_ctx* c) { return fail(c, PDEB200_EUNSUPPORTED, "Keller-Segel 2-D back-end not built"); }
int32_t kseg2d_core(pdeb200_ctx* c) { return fail(c, PDEB200_EUNSUPPORTED, "Keller-Segel 2-D back-end not built"); }
int32_t kseg2d_cost(const pdeb200_ctx*, double*, double*) { return PDEB200_EUNSUPPORTED; }
#endif
#ifndef PDEB_HAVE_NS
int32_t ns_setup(pdeb200_ctx* c) { return fail(c, PDEB200_EUNSUPPORTED, "Navier-Stokes back-end not built"); }
int32_t ns_core(pdeb200_ctx* c) { return fail(c, PDEB200_EUNSUPPORTED, "Navier-Stokes back-end not built"); }
int32_t ns_sensors(pdeb200_ctx* c, const uint8_t*) { return fail(c, PDEB200_EUNSUPPORTED, "Navier-Stokes back-end not built"); }
int32_t ns_cost(const pdeb200_ctx*, double*, double*) { return PDEB200_EUNSUPPORTED; }
void ns_free(pdeb200_ctx*) {}
#endif
#ifndef PDEB_HAVE_AGENT
void agent_free(pdeb200_ctx*) {}
double* agent_stats(pdeb200_ctx*) { return nullptr; }
#endif
}  // namespace pdeb200
#ifndef PDEB_HAVE_AGENT
extern "C" {
#define UNSUP(c) return pdeb200::fail(c, PDEB200_EUNSUPPORTED, "agent back-end not built")
int32_t pdeb200_traj_create(pdeb200_ctx* c, int64_t) { UNSUP(c); }
int32_t pdeb200_traj_length(const pdeb200_ctx* c, int64_t*) { UNSUP(c); }
int32_t pdeb200_traj_push_pre(pdeb200_ctx* c) { UNSUP(c); }
int32_t pdeb200_traj_push_post(pdeb200_ctx* c) { UNSUP(c); }
int32_t pdeb200_traj_episode_end(pdeb200_ctx* c) { UNSUP(c); }
int32_t pdeb200_traj_pop_tail(pdeb200_ctx* c) { UNSUP(c); }
int32_t pdeb200_sample(pdeb200_ctx* c, int32_t, const int64_t*, uint64_t, uint64_t) { UNSUP(c); }
int32_t pdeb200_set_batch(pdeb200_ctx* c, int32_t, const float*, const float*, const float*, const uint8_t*, const float*) { UNSUP(c); }
int32_t pdeb200_ddpg_set_path(pdeb200_ctx* c, int32_t) { UNSUP(c); }
int32_t pdeb200_ddpg_critic_grads(pdeb200_ctx* c, double, int32_t, int64_t) { UNSUP(c); }
int32_t pdeb200_ddpg_critic_apply(pdeb200_ctx* c, double) { UNSUP(c); }
int32_t pdeb200_ddpg_actor_grads(pdeb200_ctx* c, int64_t) { UNSUP(c); }
int32_t pdeb200_ddpg_actor_apply(pdeb200_ctx* c, double, double) { UNSUP(c); }
int32_t pdeb200_ddpg_update(pdeb200_ctx* c, double, double, double, double, int32_t) { UNSUP(c); }
}
#endif
