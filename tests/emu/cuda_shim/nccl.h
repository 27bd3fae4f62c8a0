// Host stand-in for nccl.h (declarations only; the library dlopens NCCL) -- TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
typedef struct shim_nccl_comm* ncclComm_t;
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
typedef enum { ncclChar = 0 } ncclDataType_t;
struct ncclUniqueId { char internal[128]; };
ncclResult_t ncclGetUniqueId(ncclUniqueId*);
ncclResult_t ncclCommInitRank(ncclComm_t*, int, ncclUniqueId, int);
ncclResult_t ncclCommDestroy(ncclComm_t);
ncclResult_t ncclGroupStart();
ncclResult_t ncclGroupEnd();
ncclResult_t ncclSend(const void*, size_t, ncclDataType_t, int, ncclComm_t, struct shim_stream*);
ncclResult_t ncclRecv(void*, size_t, ncclDataType_t, int, ncclComm_t, struct shim_stream*);
const char* ncclGetErrorString(ncclResult_t);
