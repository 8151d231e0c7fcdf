/* ucc_b200 — API version (tracks the UCC 1.9 API surface). */
#ifndef UCC_VERSION_H_
#define UCC_VERSION_H_

#define UCC_VERSION_MAJOR_SHIFT 24
#define UCC_VERSION_MINOR_SHIFT 16
#define UCC_VERSION(_maj, _min) \
    (((_maj) << UCC_VERSION_MAJOR_SHIFT) | ((_min) << UCC_VERSION_MINOR_SHIFT))

#define UCC_API_MAJOR      1
#define UCC_API_MINOR      9
#define UCC_API_VERSION    UCC_VERSION(UCC_API_MAJOR, UCC_API_MINOR)
#define UCC_VERSION_STRING "1.9.0-b200"
#ifndef UCC_GIT_REVISION
#define UCC_GIT_REVISION   "b200-native"
#endif

#endif
