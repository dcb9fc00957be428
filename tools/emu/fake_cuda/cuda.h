// fake cuda.h -- driver-API types named by make_sm_partition() in qrl_b200.cu (the emulated build never gets the entry points, so the
// library takes its plain-stream path).  TEST INFRASTRUCTURE ONLY.
#pragma once
typedef int CUdevice;
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
typedef struct emuGreenCtx* CUgreenCtx;
typedef struct emuStream* CUstream;
typedef enum { CU_DEV_RESOURCE_TYPE_SM = 1 } CUdevResourceType;
struct CUdevResource { struct { unsigned smCount; } sm; };
typedef struct emuResDesc* CUdevResourceDesc;
enum { CU_GREEN_CTX_DEFAULT_STREAM = 1, CU_STREAM_NON_BLOCKING = 1 };
