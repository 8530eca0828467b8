// stub_driver -- loads a gateway built against mex_stub (dlopen), feeds it a binary file of doubles and
// writes the outputs back; tests/test_mex_gateways.py drives it on the GPU box and compares with the
// ctypes path.  usage: stub_driver <gateway.so> <nlhs> <out.bin> <in1.bin:d0xd1[xd2]> ...
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mex.h"

typedef void (*mexfn_t)(int, mxArray**, int, const mxArray**);

static mxArray* load(const char* spec) {
    std::string s(spec);
    const size_t c = s.rfind(':');
    std::string path = s.substr(0, c), dimstr = s.substr(c + 1);
    mwSize dims[4] = {1, 1, 1, 1}; int nd = 0;
    char* tok = std::strtok(&dimstr[0], "x");
    while (tok && nd < 4) { dims[nd++] = (mwSize)std::atoll(tok); tok = std::strtok(nullptr, "x"); }
    mxArray* a = mxCreateNumericArray(nd, dims, mxDOUBLE_CLASS, mxREAL);
    size_t n = 1; for (int i = 0; i < nd; ++i) n *= dims[i];
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f || std::fread(a->data, sizeof(double), n, f) != n) { std::fprintf(stderr, "cannot read %s\n", path.c_str()); std::exit(2); }
    std::fclose(f);
    return a;
}

int main(int argc, char** argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: stub_driver gateway.so nlhs out.bin in:dims...\n"); return 2; }
    void* lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { std::fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    mexfn_t fn = (mexfn_t)dlsym(lib, "mexFunction");
    if (!fn) { std::fprintf(stderr, "no mexFunction\n"); return 2; }
    const int nlhs = std::atoi(argv[2]);
    std::vector<const mxArray*> in;
    for (int i = 4; i < argc; ++i) in.push_back(load(argv[i]));
    std::vector<mxArray*> out(nlhs < 1 ? 1 : nlhs, nullptr);
    try {
        fn(nlhs, out.data(), (int)in.size(), in.data());
    } catch (const mex_stub_error& e) {
        std::printf("MEXERROR %s|%s\n", e.id.c_str(), e.what());
        return 3;
    }
    FILE* f = std::fopen(argv[3], "wb");
    for (auto* a : out) {
        size_t n = 1; for (int i = 0; i < a->ndim; ++i) n *= a->dims[i];
        std::printf("OUT %d", a->ndim);
        for (int i = 0; i < a->ndim; ++i) std::printf(" %zu", (size_t)a->dims[i]);
        std::printf("\n");
        std::fwrite(a->data, sizeof(double), n, f);
    }
    std::fclose(f);
    return 0;
}
