// CL/cl2.hpp -- stand-in for the OpenCL C++ bindings, just enough for the reference's FPGA host
// (src/fpga/src/host_spmv_bscsr.cpp, src/fpga/src/opencl_utils.hpp) to compile and RUN IN SOFTWARE.
//
// TEST INFRASTRUCTURE ONLY.  There is no OpenCL runtime or FPGA here.  Buffers are the host pointers the
// reference hands over (it creates every buffer with CL_MEM_USE_HOST_PTR), migrations are no-ops, and
// CommandQueue::enqueueTask calls `apshim_cl_task` with the arguments recorded by Kernel::setArg -- which
// oracle/ref_fpga.cpp points at the reference's own HLS top function spmv_bscsr_top_k_main.  This is the
// same "kernel as a C function" execution Vitis sw_emu performs.
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef uint64_t cl_bitfield;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_mem_migration_flags;
typedef void *cl_event;

#define CL_SUCCESS 0
#define CL_MEM_READ_ONLY (1 << 2)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_USE_HOST_PTR (1 << 3)
#define CL_MIGRATE_MEM_OBJECT_HOST (1 << 0)
#define CL_QUEUE_PROFILING_ENABLE (1 << 1)
#define CL_QUEUE_OUT_OF_ORDER_EXEC_MODE_ENABLE (1 << 0)
#define CL_DEVICE_TYPE_ACCELERATOR (1 << 3)
#define CL_PLATFORM_NAME 0x0902
#define CL_DEVICE_NAME 0x102B
#define CL_PROFILING_COMMAND_START 0x1282
#define CL_PROFILING_COMMAND_END 0x1283

inline cl_int clGetEventProfilingInfo(cl_event, cl_uint, size_t, void *v, size_t *) {
    if (v) std::memset(v, 0, sizeof(cl_ulong));
    return CL_SUCCESS;
}

namespace cl {

struct KernelArg {
    void *ptr = nullptr;        // Buffer arguments: the host pointer
    uint64_t scalar = 0;        // scalar arguments
};

class Device {
public:
    template <int>
    std::string getInfo(cl_int *err = nullptr) const { if (err) *err = CL_SUCCESS; return "software"; }
};

class Platform {
public:
    static cl_int get(std::vector<Platform> *p) { p->clear(); return CL_SUCCESS; }
    template <int>
    std::string getInfo(cl_int *err = nullptr) const { if (err) *err = CL_SUCCESS; return "software"; }
    cl_int getDevices(cl_bitfield, std::vector<Device> *d) const { d->clear(); return CL_SUCCESS; }
};

class Context {
public:
    Context() {}
    explicit Context(const Device &, void * = nullptr, void * = nullptr, void * = nullptr, cl_int *err = nullptr) { if (err) *err = CL_SUCCESS; }
};

class Event {
public:
    static cl_int waitForEvents(const std::vector<Event> &) { return CL_SUCCESS; }
    cl_int wait() const { return CL_SUCCESS; }
    template <typename T>
    cl_int getProfilingInfo(cl_uint, T *v) const { *v = 0; return CL_SUCCESS; }
};

class Memory {
public:
    void *host = nullptr;
    size_t bytes = 0;
};

class Buffer : public Memory {
public:
    Buffer() {}
    Buffer(const Context &, cl_mem_flags, size_t size, void *host_ptr = nullptr, cl_int *err = nullptr) {
        host = host_ptr;
        bytes = size;
        if (err) *err = CL_SUCCESS;
    }
};

class Program {
public:
    typedef std::vector<std::pair<const void *, size_t>> Binaries;
    Program() {}
    Program(const Context &, const std::vector<Device> &, const Binaries &, std::vector<cl_int> * = nullptr, cl_int *err = nullptr) { if (err) *err = CL_SUCCESS; }
};

class Kernel {
public:
    std::vector<KernelArg> args;
    Kernel() {}
    Kernel(const Program &, const char *, cl_int *err = nullptr) { if (err) *err = CL_SUCCESS; }
    cl_int setArg(cl_uint i, const Buffer &b) {
        if (args.size() <= i) args.resize(i + 1);
        args[i].ptr = b.host;
        return CL_SUCCESS;
    }
    template <typename T>
    cl_int setArg(cl_uint i, const T &v) {
        if (args.size() <= i) args.resize(i + 1);
        uint64_t s = 0;
        std::memcpy(&s, &v, sizeof(T) < 8 ? sizeof(T) : 8);
        args[i].scalar = s;
        return CL_SUCCESS;
    }
};

}  // namespace cl

// Provided by the translation unit that includes the kernel (oracle/ref_fpga.cpp).
void apshim_cl_task(const std::vector<cl::KernelArg> &args);

namespace cl {

class CommandQueue {
public:
    CommandQueue() {}
    CommandQueue(const Context &, const Device &, cl_bitfield = 0, cl_int *err = nullptr) { if (err) *err = CL_SUCCESS; }
    cl_int enqueueMigrateMemObjects(const std::vector<Memory> &, cl_mem_migration_flags, const std::vector<Event> * = nullptr,
                                    Event * = nullptr) const { return CL_SUCCESS; }
    cl_int enqueueTask(const Kernel &k, const std::vector<Event> * = nullptr, Event * = nullptr) const {
        apshim_cl_task(k.args);
        return CL_SUCCESS;
    }
    cl_int finish() const { return CL_SUCCESS; }
};

}  // namespace cl
