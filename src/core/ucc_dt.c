#include "ucc_dt.h"
#include <strings.h>

const size_t ucc_dt_predefined_sizes[UCC_DT_PREDEFINED_LAST] = {
    1, 2, 4, 8, 16,  /* int8..int128 */
    1, 2, 4, 8, 16,  /* uint8..uint128 */
    2, 4, 8, 2, 16,  /* f16 f32 f64 bf16 f128 */
    8, 16, 32        /* complex f32/f64/f128 */
};
static const char *dt_names[UCC_DT_PREDEFINED_LAST] = {
    "int8", "int16", "int32", "int64", "int128", "uint8", "uint16", "uint32", "uint64", "uint128",
    "float16", "float32", "float64", "bfloat16", "float128", "float32_complex", "float64_complex", "float128_complex"};

const char *ucc_datatype_str(ucc_datatype_t dt)
{
    if (UCC_DT_IS_PREDEFINED(dt) && UCC_DT_PREDEFINED_ID(dt) < UCC_DT_PREDEFINED_LAST) return dt_names[UCC_DT_PREDEFINED_ID(dt)];
    return UCC_DT_IS_GENERIC(dt) ? "generic" : "unknown";
}
ucc_datatype_t ucc_datatype_from_str(const char *s)
{
    for (unsigned i = 0; i < UCC_DT_PREDEFINED_LAST; i++) if (!strcasecmp(s, dt_names[i])) return UCC_PREDEFINED_DT(i);
    return (ucc_datatype_t)-1;
}

UCC_EXPORT ucc_status_t ucc_dt_create_generic(const ucc_generic_dt_ops_t *ops, void *context, ucc_datatype_t *datatype_p)
{
    ucc_dt_generic_t *g;
    if (!ops || !datatype_p) return UCC_ERR_INVALID_PARAM;
    if (posix_memalign((void **)&g, 8, sizeof(*g))) return UCC_ERR_NO_MEMORY;
    g->context = context; g->ops = *ops;
    if (!(ops->mask & UCC_GENERIC_DT_OPS_FIELD_FLAGS)) g->ops.flags = 0;
    *datatype_p = ucc_dt_from_generic(g);
    return UCC_OK;
}
UCC_EXPORT void ucc_dt_destroy(ucc_datatype_t datatype)
{ if (UCC_DT_IS_GENERIC(datatype)) free(ucc_dt_to_generic(datatype)); }
