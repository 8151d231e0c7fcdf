#include <ucc/api/ucc.h>
#include "ucc_compiler_def.h"
UCC_EXPORT const char *ucc_status_string(ucc_status_t status)
{
    static __thread char unk[48];
    switch (status) {
    case UCC_OK: return "Success";
    case UCC_INPROGRESS: return "Operation in progress";
    case UCC_OPERATION_INITIALIZED: return "Operation initialized";
    case UCC_ERR_NOT_SUPPORTED: return "Operation is not supported";
    case UCC_ERR_NOT_IMPLEMENTED: return "Operation is not implemented";
    case UCC_ERR_INVALID_PARAM: return "Invalid parameter";
    case UCC_ERR_NO_MEMORY: return "Out of memory";
    case UCC_ERR_NO_RESOURCE: return "Resources are not available";
    case UCC_ERR_NO_MESSAGE: return "No message available";
    case UCC_ERR_NOT_FOUND: return "Not found";
    case UCC_ERR_TIMED_OUT: return "Operation timed out";
    case UCC_ERR_IO_ERROR: return "Input/output error";
    default: snprintf(unk, sizeof(unk), "Unknown error %d", (int)status); return unk;
    }
}
