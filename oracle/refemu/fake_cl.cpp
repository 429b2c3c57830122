// fake_cl.cpp -- a minimal in-process "OpenCL runtime" so that the reference's OWN host class
// (/root/reference/source/Lib/TLibEncoder/TEncOpenCL.cpp, compiled unmodified from where it lies)
// and its OWN kernels (/root/reference/cl/sad.cl, bodies extracted at build time by oracle/Makefile
// into oracle/_ref/gen/, never committed) run on the CPU.
//
// TEST INFRASTRUCTURE ONLY (see oracle/hmme_oracle.h).  Implements exactly the 22 entry points
// TEncOpenCL.cpp calls.  Kernels execute synchronously at clEnqueueNDRangeKernel:
//   calcSAD_AMP : one 16x16 work-group in lock-step (oracle/refemu/simt.h)
//   compareSAD  : 593 independent single-item groups, compiled as ordinary scalar C++
// calcSAD (the 425-entry AMP_ENC_SPEEDUP kernel) is inactive and broken in the reference
//  and is refused at clCreateKernel.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "CL/cl.h"
#include "simt.h"

// ---------------------------------------------------------------- runtime objects
struct _cl_platform_id { int dummy; };
struct _cl_device_id { int dummy; };
struct _cl_context { int refs; };
struct _cl_command_queue { int dummy; };
struct _cl_program { std::string src; bool built; };
struct _cl_mem { std::vector<unsigned char> bytes; };
struct KArg { std::vector<unsigned char> val; size_t localBytes; bool set; };
struct _cl_kernel { std::string name; std::vector<KArg> args; };

static _cl_platform_id g_platform;
static _cl_device_id g_device;
static const char kDeviceName[] = "lock-step CPU emulation of cl/sad.cl (oracle/_ref)";

static long g_oobReads = 0, g_oobWrites = 0, g_launches = 0;
extern "C" void refemu_stats(long* oobReads, long* oobWrites, long* launches) {
    if (oobReads) *oobReads = g_oobReads;
    if (oobWrites) *oobWrites = g_oobWrites;
    if (launches) *launches = g_launches;
}

// ---------------------------------------------------------------- the reference's kernel bodies
namespace refk {
using simt::I32;
using simt::U32;

static void calcSAD_AMP(simt::Ptr<short> block_pixel, simt::Ptr<short> area_pixel,
                        simt::Ptr<U32> sadHorizontal, simt::Ptr<U32> sadVertical, simt::Ptr<U32> sadAMP,
                        simt::Ptr<U32> tempSad, I32 iStrideCur, I32 iStrideOrg, I32 posX, I32 posY) {
    using simt::abs_diff;
#define BLOCK_WIDTH 4
#define BLOCK_HEIGHT 4
#define get_local_id(d) simt::local_id(d)
#define get_local_size(d) ((U32)simt::ctx().lsize[d])
#define barrier(f) ((void)0)
#define if(c) for (simt::MaskGuard mg_((c)); mg_.once();)
#define unsigned
#define int simt::V
#include "calcSAD_AMP.body.inc"
#undef int
#undef unsigned
#undef if
#undef barrier
#undef get_local_size
#undef get_local_id
}

static void compareSAD_item(size_t gid_, unsigned int* tempSad, unsigned int* minSad, int* Xarray, int* Yarray,
                            unsigned int* ruiCost, int uiCost, int posX, int posY) {
#define get_global_id(d) ((unsigned int)gid_)
#define barrier(f) ((void)0)
#include "compareSAD.body.inc"
#undef barrier
#undef get_global_id
}
}  // namespace refk

// ---------------------------------------------------------------- API
template <typename T> static T argval(const _cl_kernel* k, int i) { T v; memcpy(&v, k->args[i].val.data(), sizeof(T)); return v; }

extern "C" {

cl_int clGetPlatformIDs(cl_uint n, cl_platform_id* ids, cl_uint* num) {
    if (num) *num = 1;
    if (ids && n >= 1) ids[0] = &g_platform;
    return CL_SUCCESS;
}
cl_int clGetDeviceIDs(cl_platform_id, cl_device_type type, cl_uint n, cl_device_id* ids, cl_uint* num) {
    if (!(type & (CL_DEVICE_TYPE_GPU | CL_DEVICE_TYPE_DEFAULT)) && type != CL_DEVICE_TYPE_ALL) {
        if (num) *num = 0;
        return CL_DEVICE_NOT_FOUND;
    }
    if (num) *num = 1;
    if (ids && n >= 1) ids[0] = &g_device;
    return CL_SUCCESS;
}
cl_int clGetDeviceInfo(cl_device_id, cl_device_info what, size_t sz, void* out, size_t* ret) {
    if (what != CL_DEVICE_NAME) return CL_INVALID_VALUE;
    if (ret) *ret = sizeof(kDeviceName);
    if (out) { if (sz < sizeof(kDeviceName)) return CL_INVALID_VALUE; memcpy(out, kDeviceName, sizeof(kDeviceName)); }
    return CL_SUCCESS;
}
cl_context clCreateContext(const cl_context_properties*, cl_uint, const cl_device_id*,
                           void(CL_CALLBACK*)(const char*, const void*, size_t, void*), void*, cl_int* err) {
    if (err) *err = CL_SUCCESS;
    return new _cl_context{1};
}
cl_int clReleaseContext(cl_context c) { if (!c) return CL_INVALID_CONTEXT; delete c; return CL_SUCCESS; }
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int* err) {
    if (err) *err = CL_SUCCESS;                       // the device id the reference passes here dangles (App. B11): ignored
    return new _cl_command_queue{0};
}
cl_int clFlush(cl_command_queue) { return CL_SUCCESS; }
cl_int clFinish(cl_command_queue q) { return q ? CL_SUCCESS : CL_INVALID_COMMAND_QUEUE; }

cl_program clCreateProgramWithSource(cl_context, cl_uint count, const char** strs, const size_t* lens, cl_int* err) {
    _cl_program* p = new _cl_program{std::string(), false};
    for (cl_uint i = 0; i < count; ++i) p->src += lens && lens[i] ? std::string(strs[i], lens[i]) : std::string(strs[i]);
    if (err) *err = CL_SUCCESS;
    return p;
}
cl_int clBuildProgram(cl_program p, cl_uint, const cl_device_id*, const char*, void(CL_CALLBACK*)(cl_program, void*), void*) {
    // The kernels were compiled ahead of time from /root/reference/cl/sad.cl; accept only a source that
    // names both of them (the encoder passes the file given by --KernelOpenCL).
    p->built = p->src.find("calcSAD_AMP") != std::string::npos && p->src.find("compareSAD") != std::string::npos;
    return p->built ? CL_SUCCESS : CL_BUILD_PROGRAM_FAILURE;
}
cl_int clGetProgramBuildInfo(cl_program, cl_device_id, cl_program_build_info, size_t sz, void* out, size_t* ret) {
    static const char msg[] = "fake_cl: source does not contain calcSAD_AMP + compareSAD";
    if (ret) *ret = sizeof(msg);
    if (out && sz >= sizeof(msg)) memcpy(out, msg, sizeof(msg));
    return CL_SUCCESS;
}
cl_kernel clCreateKernel(cl_program p, const char* name, cl_int* err) {
    const std::string n(name ? name : "");
    if (!p || !p->built || (n != "calcSAD_AMP" && n != "compareSAD")) { if (err) *err = CL_INVALID_KERNEL_NAME; return nullptr; }
    _cl_kernel* k = new _cl_kernel{n, {}};
    k->args.resize(n == "calcSAD_AMP" ? 10 : 8);
    if (err) *err = CL_SUCCESS;
    return k;
}
cl_int clSetKernelArg(cl_kernel k, cl_uint idx, size_t sz, const void* val) {
    if (!k) return CL_INVALID_KERNEL;
    if (idx >= k->args.size()) return CL_INVALID_ARG_INDEX;
    KArg& a = k->args[idx];
    a.set = true;
    if (val) { a.val.assign((const unsigned char*)val, (const unsigned char*)val + sz); a.localBytes = 0; }
    else { a.val.clear(); a.localBytes = sz; }
    return CL_SUCCESS;
}

cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t sz, void* host, cl_int* err) {
    _cl_mem* m = new _cl_mem;
    m->bytes.assign(sz, 0xCD);                        // uninitialised device memory is not zero
    if (host) memcpy(m->bytes.data(), host, sz);
    if (err) *err = CL_SUCCESS;
    return m;
}
cl_int clReleaseMemObject(cl_mem m) { if (!m) return CL_INVALID_MEM_OBJECT; delete m; return CL_SUCCESS; }
cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem m, cl_bool, size_t off, size_t sz, const void* src, cl_uint,
                            const cl_event*, cl_event*) {
    if (!m || off + sz > m->bytes.size()) return CL_INVALID_VALUE;
    memcpy(m->bytes.data() + off, src, sz);
    return CL_SUCCESS;
}
void* clEnqueueMapBuffer(cl_command_queue, cl_mem m, cl_bool, cl_map_flags, size_t off, size_t sz, cl_uint,
                         const cl_event*, cl_event*, cl_int* err) {
    if (!m || off + sz > m->bytes.size()) { if (err) *err = CL_INVALID_VALUE; return nullptr; }
    if (err) *err = CL_SUCCESS;
    return m->bytes.data() + off;                     // the reference keeps reading after unmap (App. B8): stays valid here
}
cl_int clEnqueueUnmapMemObject(cl_command_queue, cl_mem, void*, cl_uint, const cl_event*, cl_event*) { return CL_SUCCESS; }
cl_int clEnqueueFillBuffer(cl_command_queue, cl_mem m, const void* pat, size_t psz, size_t off, size_t sz, cl_uint,
                           const cl_event*, cl_event*) {
    if (!m || !psz || off + sz > m->bytes.size()) return CL_INVALID_VALUE;
    for (size_t o = 0; o + psz <= sz; o += psz) memcpy(m->bytes.data() + off + o, pat, psz);
    return CL_SUCCESS;
}

cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel k, cl_uint dim, const size_t*, const size_t* gsz, const size_t* lsz,
                              cl_uint, const cl_event*, cl_event*) {
    if (!k) return CL_INVALID_KERNEL;
    for (const KArg& a : k->args) if (!a.set) return CL_INVALID_KERNEL_ARGS;
    ++g_launches;
    if (k->name == "calcSAD_AMP") {
        if (dim != 2 || gsz[0] != lsz[0] || gsz[1] != lsz[1] || gsz[0] * gsz[1] > (size_t)simt::MAXW) return CL_INVALID_WORK_GROUP_SIZE;
        cl_mem cur = argval<cl_mem>(k, 0), area = argval<cl_mem>(k, 1), tmp = argval<cl_mem>(k, 5);
        std::vector<simt::U32> lh(k->args[2].localBytes / 4, simt::POISON), lv(k->args[3].localBytes / 4, simt::POISON),
            la(k->args[4].localBytes / 4, simt::POISON);
        simt::begin_group((int)lsz[0], (int)lsz[1]);
        simt::ctx().oobReads = simt::ctx().oobWrites = 0;
        refk::calcSAD_AMP(simt::Ptr<short>((short*)cur->bytes.data(), cur->bytes.size() / 2),
                          simt::Ptr<short>((short*)area->bytes.data(), area->bytes.size() / 2),
                          simt::Ptr<simt::U32>(lh.data(), lh.size()), simt::Ptr<simt::U32>(lv.data(), lv.size()),
                          simt::Ptr<simt::U32>(la.data(), la.size()),
                          simt::Ptr<simt::U32>((simt::U32*)tmp->bytes.data(), tmp->bytes.size() / 4),
                          argval<int>(k, 6), argval<int>(k, 7), argval<int>(k, 8), argval<int>(k, 9));
        g_oobReads += simt::ctx().oobReads;
        g_oobWrites += simt::ctx().oobWrites;
        return CL_SUCCESS;
    }
    if (k->name == "compareSAD") {
        if (dim != 1) return CL_INVALID_WORK_DIMENSION;
        cl_mem tmp = argval<cl_mem>(k, 0), mn = argval<cl_mem>(k, 1), xs = argval<cl_mem>(k, 2), ys = argval<cl_mem>(k, 3),
               rc = argval<cl_mem>(k, 4);
        for (size_t g = 0; g < gsz[0]; ++g)
            refk::compareSAD_item(g, (unsigned*)tmp->bytes.data(), (unsigned*)mn->bytes.data(), (int*)xs->bytes.data(),
                                  (int*)ys->bytes.data(), (unsigned*)rc->bytes.data(), argval<int>(k, 5), argval<int>(k, 6),
                                  argval<int>(k, 7));
        return CL_SUCCESS;
    }
    return CL_INVALID_KERNEL;
}

}  // extern "C"
