#ifndef UCC_CL_TYPE_H_
#define UCC_CL_TYPE_H_
typedef enum { UCC_CL_BASIC, UCC_CL_HIER, UCC_CL_DOCA_UROM, UCC_CL_ALL, UCC_CL_LAST } ucc_cl_type_t;
#endif
