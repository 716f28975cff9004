// libpytv_b200.so: error state, version, sizing queries.
#include <stdarg.h>

#include <atomic>

#include "host_common.cuh"
#include "tv_path.cuh"

namespace pytvb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace pytvb

using namespace pytvb;

extern "C" {

int pytvb_version(void) { return PYTVB_VERSION; }
#ifndef PYTVB_SRC_HASH
#define PYTVB_SRC_HASH "unknown"
#endif
const char* pytvb_build_id(void) { return PYTVB_SRC_HASH; }
const char* pytvb_last_error(void) { return g_err; }
uint64_t pytvb_launch_count(void) { return (uint64_t)g_launches.load(std::memory_order_relaxed); }

int pytvb_num_components(const pytvb_problem* pb) {
    if (check_problem(pb) != PYTVB_OK) return PYTVB_ERR_ARG;
    return axes_of(pb).Nd;
}

size_t pytvb_reduce_workspace_bytes(const pytvb_problem* pb) {
    if (check_problem(pb) != PYTVB_OK) return 0;
    long long n = max_partials(pb);
    const long long nt = tile_max_blocks(pb);
    if (nt > n) n = nt;
    return (size_t)(n + REDUCE_STAGE2 + REDUCE_HEAD) * sizeof(double);
}

size_t pytvb_tv_workspace_bytes(const pytvb_problem* pb) {
    if (check_problem(pb) != PYTVB_OK) return 0;
    if (tv_uses_tile(pb)) return 256;   // single-sweep kernel: nothing goes through memory
    const size_t es = pb->dtype == PYTVB_F32 ? 4 : 8;
    // two-sweep fallback: inverse-norm field for the slab plus one plane on each side
    return (size_t)(pb->Nz + 2) * pb->M * pb->Ni * pb->Nj * es + 256;
}

}  // extern "C"
