// gs_vmm.cu — CUDA virtual memory management for the sharded (multi-GPU) layout; see gs_vmm.h.
#include "gs_vmm.h"

#include <string.h>
#include <unistd.h>

template <class T>
bool GsVmm::sym(const char* name, T* out) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    snprintf(err_, sizeof(err_), "driver entry point %s not available", name);
    return false;
  }
  *out = reinterpret_cast<T>(fn);
  return true;
}

bool GsVmm::fail(const char* what, CUresult rc) {
  const char* s = nullptr;
  if (cuGetErrorString_) cuGetErrorString_(rc, &s);
  snprintf(err_, sizeof(err_), "%s: %s (%d)", what, s ? s : "?", (int)rc);
  return false;
}

bool GsVmm::init(int device, uint32_t world, uint32_t rank, char* err, size_t err_cap) {
  device_ = device;
  world_ = world;
  rank_ = rank;
  handles_.assign(world, std::vector<CUmemGenericAllocationHandle>());
  bool ok = sym("cuGetErrorString", &cuGetErrorString_) &&
            sym("cuMemGetAllocationGranularity", &cuMemGetAllocationGranularity_) &&
            sym("cuMemAddressReserve", &cuMemAddressReserve_) && sym("cuMemAddressFree", &cuMemAddressFree_) &&
            sym("cuMemCreate", &cuMemCreate_) && sym("cuMemRelease", &cuMemRelease_) &&
            sym("cuMemMap", &cuMemMap_) && sym("cuMemUnmap", &cuMemUnmap_) &&
            sym("cuMemSetAccess", &cuMemSetAccess_) &&
            sym("cuMemExportToShareableHandle", &cuMemExportToShareableHandle_) &&
            sym("cuMemImportFromShareableHandle", &cuMemImportFromShareableHandle_);
  if (ok) {
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device_;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUresult rc = cuMemGetAllocationGranularity_(&gran_, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
    if (rc != CUDA_SUCCESS) ok = fail("cuMemGetAllocationGranularity", rc);
  }
  if (!ok && err) snprintf(err, err_cap, "%s", err_);
  return ok;
}

void* GsVmm::reserve(size_t slice_bytes, size_t planes) {
  if (slice_bytes == 0 || slice_bytes % gran_ != 0) {
    snprintf(err_, sizeof(err_), "slice of %zu bytes is not a multiple of the %zu-byte granularity", slice_bytes, gran_);
    return nullptr;
  }
  GsVmmColumn c;
  c.slice_bytes = slice_bytes;
  c.planes = planes;
  c.first_slice = n_slices_;
  CUresult rc = cuMemAddressReserve_(&c.va, slice_bytes * planes * world_, gran_, 0, 0);
  if (rc != CUDA_SUCCESS) {
    fail("cuMemAddressReserve", rc);
    return nullptr;
  }
  n_slices_ += planes;
  cols_.push_back(c);
  return reinterpret_cast<void*>(c.va);
}

bool GsVmm::map_slice(uint32_t r, size_t slice, CUmemGenericAllocationHandle h) {
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  for (const GsVmmColumn& c : cols_) {
    if (slice < c.first_slice || slice >= c.first_slice + c.planes) continue;
    const size_t p = slice - c.first_slice;
    CUdeviceptr at = c.va + (p * world_ + r) * c.slice_bytes;
    CUresult rc = cuMemMap_(at, c.slice_bytes, 0, h, 0);
    if (rc != CUDA_SUCCESS) return fail("cuMemMap", rc);
    rc = cuMemSetAccess_(at, c.slice_bytes, &acc, 1);
    if (rc != CUDA_SUCCESS) return fail("cuMemSetAccess", rc);
    return true;
  }
  snprintf(err_, sizeof(err_), "slice %zu does not exist", slice);
  return false;
}

bool GsVmm::commit() {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device_;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  handles_[rank_].assign(n_slices_, 0);
  fds_.assign(n_slices_, -1);
  for (const GsVmmColumn& c : cols_) {
    for (size_t p = 0; p < c.planes; ++p) {
      const size_t k = c.first_slice + p;
      CUmemGenericAllocationHandle h = 0;
      CUresult rc = cuMemCreate_(&h, c.slice_bytes, &prop, 0);
      if (rc != CUDA_SUCCESS) return fail("cuMemCreate", rc);
      handles_[rank_][k] = h;
      int fd = -1;
      rc = cuMemExportToShareableHandle_(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
      if (rc != CUDA_SUCCESS) return fail("cuMemExportToShareableHandle", rc);
      fds_[k] = fd;
      if (!map_slice(rank_, k, h)) return false;
    }
  }
  return true;
}

bool GsVmm::attach(uint32_t peer, const int* fds, size_t n) {
  if (peer >= world_ || peer == rank_ || !handles_[peer].empty() || n != n_slices_) {
    snprintf(err_, sizeof(err_), "bad attach: peer %u, %zu descriptors (expected %zu)", peer, n, n_slices_);
    return false;
  }
  handles_[peer].assign(n_slices_, 0);
  for (size_t k = 0; k < n; ++k) {
    CUmemGenericAllocationHandle h = 0;
    CUresult rc = cuMemImportFromShareableHandle_(&h, (void*)(uintptr_t)fds[k], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (rc != CUDA_SUCCESS) return fail("cuMemImportFromShareableHandle", rc);
    handles_[peer][k] = h;
    if (!map_slice(peer, k, h)) return false;
  }
  return true;
}

void GsVmm::destroy() {
  for (const GsVmmColumn& c : cols_) {
    for (uint32_t r = 0; r < world_; ++r) {
      if (handles_[r].empty()) continue;
      for (size_t p = 0; p < c.planes; ++p)
        if (handles_[r][c.first_slice + p]) cuMemUnmap_(c.va + (p * world_ + r) * c.slice_bytes, c.slice_bytes);
    }
    cuMemAddressFree_(c.va, c.slice_bytes * c.planes * world_);
  }
  cols_.clear();
  for (auto& hv : handles_) {
    for (auto& h : hv)
      if (h) cuMemRelease_(h);
    hv.clear();
  }
  for (int& fd : fds_)
    if (fd >= 0) close(fd);
  fds_.clear();
}
