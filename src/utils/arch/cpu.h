/* CPU vendor/model detection used by ucc.conf section predicates. */
#ifndef UCC_ARCH_CPU_H_
#define UCC_ARCH_CPU_H_
typedef enum { UCC_CPU_VENDOR_UNKNOWN, UCC_CPU_VENDOR_INTEL, UCC_CPU_VENDOR_AMD, UCC_CPU_VENDOR_ARM, UCC_CPU_VENDOR_NVIDIA, UCC_CPU_VENDOR_LAST } ucc_cpu_vendor_t;
ucc_cpu_vendor_t ucc_arch_get_cpu_vendor(void);
const char *ucc_cpu_vendor_string(ucc_cpu_vendor_t v);
const char *ucc_arch_get_cpu_model_string(void); /* lower-case coarse model: "skylake", "rome", "grace", ... or "unknown" */
#endif
