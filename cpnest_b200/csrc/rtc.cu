// NVRTC path for user-written functors: the kernels of mh_kernel.cuh / slice_kernel.cuh / hmc_kernel.cuh /
// eval_kernel.cuh are instantiated for cpn::UserModel (user_model.cuh) at run time, for sm_100a, and launched through
// the driver API.  libnvrtc and libcuda are opened with dlopen so that the library itself still loads on a machine
// without a driver (the CPU-side ABI tests).
#include "rtc.h"

#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

namespace cpn {

namespace {
typedef int nvrtcResult;
typedef void* nvrtcProgram;
typedef int CUresult;
typedef void* CUmodule;
typedef void* CUfunction;

struct Libs {
    void *nvrtc = nullptr, *cuda = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    nvrtcResult (*DestroyProgram)(nvrtcProgram*);
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**);
    const char* (*GetErrorString)(nvrtcResult);
    CUresult (*ModuleLoadData)(CUmodule*, const void*);
    CUresult (*ModuleUnload)(CUmodule);
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**);
    CUresult (*FuncSetAttribute)(CUfunction, int, int);
    CUresult (*GetErrorStringCU)(CUresult, const char**);
    bool ok = false;
    std::string why;
};

template <class T>
bool sym(void* lib, const char* name, T& fn, std::string& why) {
    fn = (T)dlsym(lib, name);
    if (!fn) { why = std::string("missing symbol ") + name; return false; }
    return true;
}

Libs& libs() {
    static Libs L;
    static bool tried = false;
    if (tried) return L;
    tried = true;
    const char* nv[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* n : nv)
        if ((L.nvrtc = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!L.nvrtc) { L.why = "libnvrtc not found (needed to compile a user functor)"; return L; }
    L.cuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!L.cuda) { L.why = "libcuda.so.1 not found (no NVIDIA driver)"; return L; }
    bool ok = sym(L.nvrtc, "nvrtcCreateProgram", L.CreateProgram, L.why) && sym(L.nvrtc, "nvrtcDestroyProgram", L.DestroyProgram, L.why) &&
              sym(L.nvrtc, "nvrtcAddNameExpression", L.AddNameExpression, L.why) && sym(L.nvrtc, "nvrtcCompileProgram", L.CompileProgram, L.why) &&
              sym(L.nvrtc, "nvrtcGetProgramLogSize", L.GetProgramLogSize, L.why) && sym(L.nvrtc, "nvrtcGetProgramLog", L.GetProgramLog, L.why) &&
              sym(L.nvrtc, "nvrtcGetCUBINSize", L.GetCUBINSize, L.why) && sym(L.nvrtc, "nvrtcGetCUBIN", L.GetCUBIN, L.why) &&
              sym(L.nvrtc, "nvrtcGetLoweredName", L.GetLoweredName, L.why) && sym(L.nvrtc, "nvrtcGetErrorString", L.GetErrorString, L.why) &&
              sym(L.cuda, "cuModuleLoadData", L.ModuleLoadData, L.why) && sym(L.cuda, "cuModuleUnload", L.ModuleUnload, L.why) &&
              sym(L.cuda, "cuModuleGetFunction", L.ModuleGetFunction, L.why) && sym(L.cuda, "cuLaunchKernel", L.LaunchKernel, L.why) &&
              sym(L.cuda, "cuFuncSetAttribute", L.FuncSetAttribute, L.why) && sym(L.cuda, "cuGetErrorString", L.GetErrorStringCU, L.why);
    L.ok = ok;
    return L;
}

std::string expr_for(const char* kind, int G, int E) {
    char buf[160];
    if (!strcmp(kind, "eval")) snprintf(buf, sizeof buf, "cpn::k_eval_points<%d, %d, cpn::UserModel>", G, E);
    else snprintf(buf, sizeof buf, "cpn::k_%s_chains<%d, %d, cpn::UserModel, false>", kind, G, E);
    return buf;
}
}  // namespace

struct RtcModule {
    CUmodule mod = nullptr;
    std::map<std::string, CUfunction> fn;     // name expression -> function
};

RtcModule* rtc_compile(const char* user_source, const char* const* include_dirs, int n_dirs, const int (*cfgs)[2], int n_cfgs,
                       int kinds, char* err, size_t err_len) {
    Libs& L = libs();
    if (!L.ok) { snprintf(err, err_len, "%s", L.why.c_str()); return nullptr; }
    std::string src;
    src += "#include \"eval_kernel.cuh\"\n#include \"mh_kernel.cuh\"\n#include \"slice_kernel.cuh\"\n#include \"hmc_kernel.cuh\"\n";
    src += "#line 1 \"user_functor.cu\"\n";
    src += user_source;
    src += "\n#include \"user_model.cuh\"\n";
    nvrtcProgram prog;
    nvrtcResult r = L.CreateProgram(&prog, src.c_str(), "cpnest_b200_user.cu", 0, nullptr, nullptr);
    if (r != 0) { snprintf(err, err_len, "nvrtcCreateProgram: %s", L.GetErrorString(r)); return nullptr; }
    static const char* kKinds[4] = {"eval", "mh", "slice", "hmc"};
    std::vector<std::string> exprs;
    for (int k = 0; k < 4; ++k)
        if (kinds & (1 << k))
            for (int i = 0; i < n_cfgs; ++i) exprs.push_back(expr_for(kKinds[k], cfgs[i][0], cfgs[i][1]));
    for (auto& e : exprs) L.AddNameExpression(prog, e.c_str());
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
    for (int i = 0; i < n_dirs; ++i) opts.push_back(std::string("-I") + include_dirs[i]);
    std::vector<const char*> copts;
    for (auto& o : opts) copts.push_back(o.c_str());
    r = L.CompileProgram(prog, (int)copts.size(), copts.data());
    if (r != 0) {
        size_t n = 0;
        L.GetProgramLogSize(prog, &n);
        std::string log(n + 1, '\0');
        if (n) L.GetProgramLog(prog, &log[0]);
        snprintf(err, err_len, "compiling the user functor failed (%s):\n%s", L.GetErrorString(r), log.c_str());
        L.DestroyProgram(&prog);
        return nullptr;
    }
    size_t nb = 0;
    L.GetCUBINSize(prog, &nb);
    std::vector<char> cubin(nb);
    L.GetCUBIN(prog, cubin.data());
    RtcModule* m = new RtcModule();
    CUresult cr = L.ModuleLoadData(&m->mod, cubin.data());
    if (cr != 0) {
        const char* s = nullptr;
        L.GetErrorStringCU(cr, &s);
        snprintf(err, err_len, "cuModuleLoadData: %s", s ? s : "?");
        L.DestroyProgram(&prog);
        delete m;
        return nullptr;
    }
    for (auto& e : exprs) {
        const char* low = nullptr;
        if (L.GetLoweredName(prog, e.c_str(), &low) != 0 || !low) { snprintf(err, err_len, "no lowered name for %s", e.c_str()); continue; }
        CUfunction f = nullptr;
        if (L.ModuleGetFunction(&f, m->mod, low) == 0) m->fn[e] = f;
    }
    L.DestroyProgram(&prog);
    return m;
}

void rtc_destroy(RtcModule* m) {
    if (!m) return;
    if (m->mod) libs().ModuleUnload(m->mod);
    delete m;
}

int rtc_launch(RtcModule* m, const char* kind, int G, int E, const void* params, int nblocks, int nthreads, size_t smem,
               cudaStream_t stream, char* err, size_t err_len) {
    if (!m) { snprintf(err, err_len, "user functor not compiled: call cpn_set_user_functor first"); return 1; }
    auto it = m->fn.find(expr_for(kind, G, E));
    if (it == m->fn.end()) { snprintf(err, err_len, "user functor: kernel %s<%d,%d> was not compiled", kind, G, E); return 1; }
    Libs& L = libs();
    if (smem > 48 * 1024) L.FuncSetAttribute(it->second, 8 /*CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES*/, (int)smem);
    void* args[] = {const_cast<void*>(params)};
    CUresult cr = L.LaunchKernel(it->second, nblocks, 1, 1, nthreads, 1, 1, (unsigned)smem, (void*)stream, args, nullptr);
    if (cr != 0) {
        const char* s = nullptr;
        L.GetErrorStringCU(cr, &s);
        snprintf(err, err_len, "cuLaunchKernel(%s<%d,%d>): %s", kind, G, E, s ? s : "?");
        return 1;
    }
    return 0;
}

}  // namespace cpn
