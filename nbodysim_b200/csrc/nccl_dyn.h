// nccl_dyn.h -- NCCL bound at run time (dlopen), so the library has no link-time dependency:
// inside a torch process `libnccl.so.2` resolves to the copy torch already loaded (2.28.x); in
// the stand-alone C driver it resolves to the system library.  Prototypes restated from nccl.h.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstddef>

namespace nb {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_UINT8 = 1, NCCL_INT32 = 2, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

struct Nccl {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;

    bool load()
    {
        if (handle) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) return false;
#define NB_SYM(field, name) \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name)); \
    if (!field) return false;
        NB_SYM(GetUniqueId, "ncclGetUniqueId")
        NB_SYM(CommInitRank, "ncclCommInitRank")
        NB_SYM(CommInitAll, "ncclCommInitAll")
        NB_SYM(CommDestroy, "ncclCommDestroy")
        NB_SYM(AllGather, "ncclAllGather")
        NB_SYM(AllReduce, "ncclAllReduce")
        NB_SYM(GroupStart, "ncclGroupStart")
        NB_SYM(GroupEnd, "ncclGroupEnd")
        NB_SYM(GetErrorString, "ncclGetErrorString")
#undef NB_SYM
        return true;
    }
};

inline Nccl &nccl()
{
    static Nccl n;
    return n;
}

} // namespace nb
