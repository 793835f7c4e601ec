// TEST INFRASTRUCTURE ONLY (oracle build shim) -- never included by the product.
//
// Stand-in for the Xilinx "xcl2.hpp" + OpenCL C++ bindings that the reference
// module headers include.  Every cl:: object is an inert value type: the
// oracle only calls the reference's compute_reference_results() functions,
// which never touch the device.  The aligned allocator keeps the reference's
// contract (4096-byte aligned host vectors).
#ifndef ORACLE_SHIM_XCL2_HPP_
#define ORACLE_SHIM_XCL2_HPP_

#include <algorithm>
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <new>
#include <string>
#include <utility>
#include <vector>

typedef int cl_int;
typedef unsigned long long cl_mem_flags;
struct cl_mem_ext_ptr_t { unsigned flags; void *obj; void *param; };

#define CL_SUCCESS 0
#define CL_DEVICE_NAME 0x102B
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_ONLY (1 << 2)
#define CL_MEM_USE_HOST_PTR (1 << 3)
#define CL_MEM_EXT_PTR_XILINX (1u << 31)
#define CL_MIGRATE_MEM_OBJECT_HOST (1 << 0)
#define CL_QUEUE_OUT_OF_ORDER_EXEC_MODE_ENABLE (1 << 0)
#define CL_QUEUE_PROFILING_ENABLE (1 << 1)
#define XCL_MEM_TOPOLOGY (1u << 31)

#define OCL_CHECK(error, call)                                                   \
    call;                                                                        \
    if (error != CL_SUCCESS) {                                                   \
        printf("%s:%d Error calling " #call ", error code is: %d\n", __FILE__,   \
               __LINE__, error);                                                 \
        exit(EXIT_FAILURE);                                                      \
    }

template <typename T>
struct aligned_allocator {
    using value_type = T;
    aligned_allocator() {}
    template <typename U> aligned_allocator(const aligned_allocator<U> &) {}
    T *allocate(std::size_t num) {
        void *ptr = nullptr;
        if (posix_memalign(&ptr, 4096, (num ? num : 1) * sizeof(T))) throw std::bad_alloc();
        return reinterpret_cast<T *>(ptr);
    }
    void deallocate(T *p, std::size_t) { free(p); }
    template <typename U> bool operator==(const aligned_allocator<U> &) const { return true; }
    template <typename U> bool operator!=(const aligned_allocator<U> &) const { return false; }
};

namespace cl {

struct Device {
    Device() {}
    Device(std::nullptr_t) {}
    template <int Name> std::string getInfo() const { return std::string("oracle-shim"); }
};

struct Context {
    Context() {}
    Context(std::nullptr_t) {}
    Context(const Device &, void *, void *, void *) {}
};

struct Buffer {
    Buffer() {}
    Buffer(const Context &, cl_mem_flags, std::size_t, void * = nullptr, cl_int *err = nullptr) {
        if (err) *err = CL_SUCCESS;
    }
};

struct Program {
    typedef std::vector<std::pair<const void *, std::size_t>> Binaries;
    Program() {}
    Program(const Context &, const std::vector<Device> &, const Binaries &, void * = nullptr,
            cl_int *err = nullptr) {
        if (err) *err = CL_SUCCESS;
    }
};

struct Kernel {
    Kernel() {}
    Kernel(std::nullptr_t) {}
    Kernel(const Program &, const char *, cl_int *err = nullptr) {
        if (err) *err = CL_SUCCESS;
    }
    template <typename T> cl_int setArg(unsigned, const T &) { return CL_SUCCESS; }
    cl_int setArg(unsigned, std::size_t, const void *) { return CL_SUCCESS; }
};

struct CommandQueue {
    CommandQueue() {}
    CommandQueue(std::nullptr_t) {}
    CommandQueue(const Context &, const Device &, cl_mem_flags = 0, cl_int *err = nullptr) {
        if (err) *err = CL_SUCCESS;
    }
    cl_int enqueueCopyBuffer(const Buffer &, const Buffer &, std::size_t, std::size_t, std::size_t) {
        return CL_SUCCESS;
    }
    cl_int enqueueMigrateMemObjects(const std::vector<Buffer> &, cl_mem_flags) { return CL_SUCCESS; }
    cl_int enqueueTask(const Kernel &) { return CL_SUCCESS; }
    cl_int finish() { return CL_SUCCESS; }
};

}  // namespace cl

namespace xcl {
inline std::vector<cl::Device> get_xil_devices() { return std::vector<cl::Device>(1); }
inline std::vector<unsigned char> read_binary_file(const std::string &) {
    return std::vector<unsigned char>(4, 0);
}
}  // namespace xcl

#endif  // ORACLE_SHIM_XCL2_HPP_
