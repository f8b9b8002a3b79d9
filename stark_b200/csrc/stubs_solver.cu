// Temporary: assembly / solver entry points until assembly.cu, pcg.cu, newton.cu land.
#include "internal.h"
namespace sb { void assembly_destroy(sb_context*) {} void pcg_destroy(sb_context*) {} }
using namespace sb;
extern "C" {
#define NOT_YET(name) return fail(ctx, SB_ERR_STATE, name ": not built yet")
int sb_project_to_pd(sb_context* ctx, double, double, int, int64_t*, int64_t*, int*) { NOT_YET("sb_project_to_pd"); }
int sb_assemble(sb_context* ctx) { NOT_YET("sb_assemble"); }
int sb_bcsr_info(sb_context* ctx, int*, int64_t*) { NOT_YET("sb_bcsr_info"); }
int sb_bcsr_get(sb_context* ctx, int64_t*, int32_t*, float*) { NOT_YET("sb_bcsr_get"); }
int sb_solve_pcg(sb_context* ctx, double, double, int, int, int*, int*, double*, double*) { NOT_YET("sb_solve_pcg"); }
void sb_newton_default_settings(sb_newton_settings*) {}
int sb_newton_solve(sb_context* ctx, const sb_newton_settings*, sb_newton_stats*) { NOT_YET("sb_newton_solve"); }
}
