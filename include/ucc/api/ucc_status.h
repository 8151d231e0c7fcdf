/* ucc_b200 — status codes of the UCC-compatible C API.
 * Values match the contract documented in reference src/ucc/api/ucc_status.h:25-45. */
#ifndef UCC_STATUS_H_
#define UCC_STATUS_H_

#ifdef __cplusplus
#define BEGIN_C_DECLS extern "C" {
#define END_C_DECLS   }
#else
#define BEGIN_C_DECLS
#define END_C_DECLS
#endif

BEGIN_C_DECLS

typedef enum {
    /* success / progress states (non-negative) */
    UCC_OK                    = 0,
    UCC_INPROGRESS            = 1,
    UCC_OPERATION_INITIALIZED = 2,
    /* failures (negative) */
    UCC_ERR_NOT_SUPPORTED   = -1,
    UCC_ERR_NOT_IMPLEMENTED = -2,
    UCC_ERR_INVALID_PARAM   = -3,
    UCC_ERR_NO_MEMORY       = -4,
    UCC_ERR_NO_RESOURCE     = -5,
    UCC_ERR_NO_MESSAGE      = -6,
    UCC_ERR_NOT_FOUND       = -7,
    UCC_ERR_TIMED_OUT       = -8,
    UCC_ERR_IO_ERROR        = -9,
    UCC_ERR_LAST            = -100
} ucc_status_t;

/* Human-readable name of a status code; never returns NULL. */
const char *ucc_status_string(ucc_status_t status);

END_C_DECLS
#endif
