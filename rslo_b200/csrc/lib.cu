// Library-level entry points: ABI version and last-error text.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace rslo {
static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;
void set_last_error(const char* what, cudaError_t e)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
}
static bool g_capture_hint = false;
bool pdl_enabled()
{
    static const bool on = [] {
        const char* e = getenv("RSLO_PDL");
        return !(e && e[0] == '0');
    }();
    return on && g_capture_hint;
}
void set_capture_hint(bool on) { g_capture_hint = on; }
}  // namespace rslo

extern "C" int rslo_abi_version(void) { return 5; }   // 5: rslo_set_graph_capture_hint; 4: optimizer step (optim.cu); 3: dense head (conv2d_tc, head_ops); 2: rslo_kabsch gained tgt_idx + normal; tensor-core entry points
extern "C" const char* rslo_last_error(void) { return rslo::g_err; }

extern "C" unsigned long long rslo_kernel_launch_count(void) { return rslo::g_launch_count; }
extern "C" void rslo_set_graph_capture_hint(int on) { rslo::set_capture_hint(on != 0); }
