/* Lightweight profiling (UCS-free): accumulate or log timestamps per named
 * location.  Macro slots mirror reference utils/profile/ucc_profile_on.h:34-96;
 * enabled at build time per component with -DUCC_PROFILING_<COMP>. */
#ifndef UCC_PROFILE_H_
#define UCC_PROFILE_H_
#include "utils/ucc_compiler_def.h"
#include "utils/ucc_time.h"

typedef struct ucc_profile_loc { const char *name; const char *file; int line; uint64_t count; double total; int registered; } ucc_profile_loc_t;
void ucc_profile_init(const char *mode_str, const char *file, size_t log_size);
void ucc_profile_cleanup(void);
void ucc_profile_record(ucc_profile_loc_t *loc, double t_begin, double t_end, const void *req);
extern unsigned ucc_profile_mode_mask;

#define UCC_PROFILE_SCOPE_BEGIN_(_name) \
    static ucc_profile_loc_t _ucc_prof_loc = {_name, __FILE__, __LINE__, 0, 0.0, 0}; \
    double _ucc_prof_t0 = ucc_profile_mode_mask ? ucc_get_time() : 0.0
#define UCC_PROFILE_SCOPE_END_(_req) \
    do { if (ucc_unlikely(ucc_profile_mode_mask)) ucc_profile_record(&_ucc_prof_loc, _ucc_prof_t0, ucc_get_time(), _req); } while (0)
#define UCC_PROFILE_EVENT_(_name, _req) \
    do { if (ucc_unlikely(ucc_profile_mode_mask)) { static ucc_profile_loc_t _l = {_name, __FILE__, __LINE__, 0, 0.0, 0}; \
         double _t = ucc_get_time(); ucc_profile_record(&_l, _t, _t, _req); } } while (0)

#define UCC_PROFILE_FUNC_BEGIN(_name)        UCC_PROFILE_SCOPE_BEGIN_(_name)
#define UCC_PROFILE_FUNC_END()               UCC_PROFILE_SCOPE_END_(NULL)
#define UCC_PROFILE_REQUEST_NEW(_req, _name, _param)   UCC_PROFILE_EVENT_(_name, _req)
#define UCC_PROFILE_REQUEST_EVENT(_req, _name, _param) UCC_PROFILE_EVENT_(_name, _req)
#define UCC_PROFILE_REQUEST_FREE(_req)                 UCC_PROFILE_EVENT_("request_free", _req)

/* event whose name is composed at run time ("<prefix>_<coll>_<suffix>", reference e.g. "ucp_allreduce_kn_start"): the location is
 * looked up by name, so this is slower than the static macros - only called when profiling is on */
void ucc_profile_event_named(const char *prefix, const char *mid, const char *suffix, const void *req);
#define UCC_PROFILE_REQUEST_EVENT_NAMED(_req, _prefix, _mid, _suffix) \
    do { if (ucc_unlikely(ucc_profile_mode_mask)) ucc_profile_event_named(_prefix, _mid, _suffix, _req); } while (0)

/* NVTX-style range hooks used around kernel launches; resolved at runtime if libnvToolsExt is loadable */
void ucc_profile_range_push(const char *name);
void ucc_profile_range_pop(void);
#endif
