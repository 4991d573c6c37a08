// wl_dist.h — z-slab multi-GPU plumbing: one process per GPU, NCCL over NVLink for halo planes, scalar all-reduces and the
// gather of coarse multigrid levels.  NCCL is dlopen'ed on first use so that single-GPU users carry no dependency and so that a
// host process that already loaded an NCCL (torch does) shares that copy instead of clashing with a second one.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
enum { WL_NCCL_SUM = 0, WL_NCCL_MAX = 2, WL_NCCL_FLOAT = 7, WL_NCCL_DOUBLE = 8 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(const char** why) {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      *why = "dlopen(libnccl.so.2) failed";
      return false;
    }
#define WL_SYM(field, name)                       \
  *(void**)(&field) = dlsym(lib, name);           \
  if (!field) {                                   \
    *why = "missing NCCL symbol " name;           \
    return false;                                 \
  }
    WL_SYM(GetUniqueId, "ncclGetUniqueId")
    WL_SYM(CommInitRank, "ncclCommInitRank")
    WL_SYM(CommDestroy, "ncclCommDestroy")
    WL_SYM(Send, "ncclSend")
    WL_SYM(Recv, "ncclRecv")
    WL_SYM(GroupStart, "ncclGroupStart")
    WL_SYM(GroupEnd, "ncclGroupEnd")
    WL_SYM(AllReduce, "ncclAllReduce")
    WL_SYM(AllGather, "ncclAllGather")
    WL_SYM(GetErrorString, "ncclGetErrorString")
#undef WL_SYM
    return true;
  }
};

struct Dist {
  int rank = 0, P = 1;
  ncclComm_t comm = nullptr;
  int up = -1, down = -1;  // ranks owning the slab above / below (−1: wall)
  bool on() const { return P > 1; }
};
