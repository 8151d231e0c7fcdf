/* Datatype helpers: predefined size table + generic (user) datatypes.
 * Reference behaviour: core/ucc_dt.c:10-57 (sizes), ucc.h:226-433. */
#ifndef UCC_DT_H_
#define UCC_DT_H_
#include <ucc/api/ucc.h>
#include "utils/ucc_compiler_def.h"

struct ucc_dt_generic {
    void                *context;
    ucc_generic_dt_ops_t ops;
};

extern const size_t ucc_dt_predefined_sizes[UCC_DT_PREDEFINED_LAST];

#define UCC_DT_IS_PREDEFINED(_dt) (((_dt) & UCC_DATATYPE_CLASS_MASK) == UCC_DATATYPE_PREDEFINED)
#define UCC_DT_IS_GENERIC(_dt)    (((_dt) & UCC_DATATYPE_CLASS_MASK) == UCC_DATATYPE_GENERIC)
#define UCC_DT_PREDEFINED_ID(_dt) ((unsigned)((_dt) >> UCC_DATATYPE_SHIFT))

static inline ucc_dt_generic_t *ucc_dt_to_generic(ucc_datatype_t dt)
{ return (ucc_dt_generic_t *)(void *)(uintptr_t)(dt & ~(uint64_t)UCC_DATATYPE_CLASS_MASK); }
static inline ucc_datatype_t ucc_dt_from_generic(ucc_dt_generic_t *g)
{ return ((uint64_t)(uintptr_t)g) | UCC_DATATYPE_GENERIC; }
static inline int UCC_DT_IS_CONTIG(ucc_datatype_t dt)
{ return UCC_DT_IS_PREDEFINED(dt) || (UCC_DT_IS_GENERIC(dt) && (ucc_dt_to_generic(dt)->ops.flags & UCC_GENERIC_DT_OPS_FLAG_CONTIG)); }
static inline int UCC_DT_HAS_REDUCE(ucc_datatype_t dt)
{ return UCC_DT_IS_GENERIC(dt) && (ucc_dt_to_generic(dt)->ops.flags & UCC_GENERIC_DT_OPS_FLAG_REDUCE); }
static inline size_t ucc_dt_size(ucc_datatype_t dt)
{
    if (UCC_DT_IS_PREDEFINED(dt)) { unsigned id = UCC_DT_PREDEFINED_ID(dt); return id < UCC_DT_PREDEFINED_LAST ? ucc_dt_predefined_sizes[id] : 0; }
    if (UCC_DT_IS_CONTIG(dt)) return ucc_dt_to_generic(dt)->ops.contig_size;
    return 0;
}
const char *ucc_datatype_str(ucc_datatype_t dt);
ucc_datatype_t ucc_datatype_from_str(const char *s); /* (ucc_datatype_t)-1 on error */
#endif
