// nccl_dyn.hpp -- NCCL entry points resolved at run time (dlopen), so that the library loads on machines without NCCL
// and, inside a PyTorch process, binds to the NCCL that torch.distributed already loaded (same soname).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace lbm {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                                        = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                                 = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t)                                                           = nullptr;
  ncclResult_t (*GroupStart)()                                                                      = nullptr;
  ncclResult_t (*GroupEnd)()                                                                        = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)          = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)                = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t)                                                       = nullptr;

  bool load(std::string* err) {
    if(handle != nullptr) return true;
    for(const char* name : {"libnccl.so.2", "libnccl.so"}) {
      handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if(handle != nullptr) break;
    }
    if(handle == nullptr) { *err = std::string("cannot load NCCL: ") + dlerror(); return false; }
    auto sym = [&](const char* n) { return dlsym(handle, n); };
    GetUniqueId    = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
    CommInitRank   = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
    CommDestroy    = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    GroupStart     = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
    GroupEnd       = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
    Send           = reinterpret_cast<decltype(Send)>(sym("ncclSend"));
    Recv           = reinterpret_cast<decltype(Recv)>(sym("ncclRecv"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    AllReduce      = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
    if(!AllReduce || !GetUniqueId || !CommInitRank || !CommDestroy || !GroupStart || !GroupEnd || !Send || !Recv || !GetErrorString) {
      *err = "NCCL library lacks a required symbol";
      return false;
    }
    return true;
  }
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  return api;
}

} // namespace lbm
