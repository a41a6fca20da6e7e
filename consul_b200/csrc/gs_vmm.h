// gs_vmm.h — one virtual address range per column, physically sharded across the GPUs of a box.
//
// Multi-GPU layout: rank r owns the members [r*S, (r+1)*S).  Every column is a
// CUDA virtual-memory reservation of planes * world * slice bytes; the slice (plane p, rank r)
// is backed by rank r's HBM and mapped at the same offset in every process, so device code
// indexes a column by global member id exactly as on one GPU and the hardware routes the access
// to local HBM or over NVLink.  cuMemMap cannot map a sub-range of a physical allocation
// (its offset argument must be zero), so every (column, plane) slice is its own physical
// allocation: each rank exports one POSIX file descriptor per slice and maps the other ranks'
// slices after importing their descriptors.  The driver API is reached through cudaGetDriverEntryPoint so
// libgsim.so has no link-time dependency on libcuda (it must still load on GPU-less hosts).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

struct GsVmmColumn {
  CUdeviceptr va = 0;
  size_t slice_bytes = 0;  // bytes of one (plane, rank) slice: a multiple of the granularity
  size_t planes = 0;
  size_t first_slice = 0;  // index of (plane 0) in the per-rank slice list
};

class GsVmm {
 public:
  bool init(int device, uint32_t world, uint32_t rank, char* err, size_t err_cap);
  size_t granularity() const { return gran_; }
  // reserve the address range of one column (all planes, all ranks); returns its base pointer
  void* reserve(size_t slice_bytes, size_t planes);
  // create and map this rank's slices of every column reserved so far
  bool commit();
  const std::vector<int>& export_fds() const { return fds_; }
  bool attach(uint32_t peer, const int* fds, size_t n);  // import and map a peer's slices
  void destroy();
  const char* last_error() const { return err_; }

 private:
  bool map_slice(uint32_t r, size_t slice, CUmemGenericAllocationHandle h);
  bool fail(const char* what, CUresult rc);
  template <class T>
  bool sym(const char* name, T* out);

  int device_ = 0;
  uint32_t world_ = 1, rank_ = 0;
  size_t gran_ = 0, n_slices_ = 0;
  std::vector<int> fds_;  // this rank's exported descriptors, one per slice
  std::vector<GsVmmColumn> cols_;
  std::vector<std::vector<CUmemGenericAllocationHandle>> handles_;  // [world][slice], 0 = not mapped
  char err_[256] = {0};

  // driver entry points
  CUresult (*cuMemGetAllocationGranularity_)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*cuMemAddressReserve_)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*cuMemAddressFree_)(CUdeviceptr, size_t) = nullptr;
  CUresult (*cuMemCreate_)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*cuMemRelease_)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*cuMemMap_)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*cuMemUnmap_)(CUdeviceptr, size_t) = nullptr;
  CUresult (*cuMemSetAccess_)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*cuMemExportToShareableHandle_)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*cuMemImportFromShareableHandle_)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*cuGetErrorString_)(CUresult, const char**) = nullptr;
};
