// Temporary: contact entry points until contact.cu lands.
#include "internal.h"
namespace sb { void contact_destroy(sb_context*) {} int contact_update_internal(sb_context*) { return 0; } int contact_intersections_internal(sb_context*, int* c) { *c = 0; return 0; } bool contact_active(sb_context*) { return false; } }
using namespace sb;
extern "C" {
#define NOT_YET(name) return fail(ctx, SB_ERR_STATE, name ": not built yet")
int sb_contact_init(sb_context* ctx, const sb_contact_bindings*) { NOT_YET("sb_contact_init"); }
int sb_contact_add_mesh(sb_context* ctx, const sb_contact_mesh*, int*) { NOT_YET("sb_contact_add_mesh"); }
int sb_contact_blacklist(sb_context* ctx, int, int) { NOT_YET("sb_contact_blacklist"); }
int sb_contact_set_friction(sb_context* ctx, int, int, double) { NOT_YET("sb_contact_set_friction"); }
int sb_contact_set_params(sb_context* ctx, double, double, int, int, int) { NOT_YET("sb_contact_set_params"); }
int sb_contact_update(sb_context* ctx) { NOT_YET("sb_contact_update"); }
int sb_contact_update_friction(sb_context* ctx) { NOT_YET("sb_contact_update_friction"); }
int sb_contact_count_intersections(sb_context* ctx, int*) { NOT_YET("sb_contact_count_intersections"); }
int sb_contact_get_proximity(sb_context* ctx, int, int32_t*, double*, int, int*, int*) { NOT_YET("sb_contact_get_proximity"); }
int sb_contact_get_vertices(sb_context* ctx, int, double*) { NOT_YET("sb_contact_get_vertices"); }
int sb_contact_set_vertices(sb_context* ctx, int, const double*) { NOT_YET("sb_contact_set_vertices"); }
int sb_contact_detect(sb_context* ctx, double, int) { NOT_YET("sb_contact_detect"); }
int sb_contact_potential(sb_context* ctx, const char*, int*) { NOT_YET("sb_contact_potential"); }
}
