// Per-kernel device timing for bench.py: when enabled, every kernel launch of the library is
// bracketed by a pair of CUDA events recorded on the launching stream; ttl_prof_report sums
// the elapsed times per kernel name.  Off by default (zero overhead beyond one branch).
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "ttl_common.cuh"

std::atomic<int> g_ttl_pdl{1};

namespace {
struct Pending { const char* name; cudaEvent_t a, b; };
bool g_on = false;
std::vector<Pending> g_pending;
std::vector<cudaEvent_t> g_pool;
const char* g_cur_name = nullptr;
cudaEvent_t g_cur_a = nullptr;

cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void ttl_prof_begin(const char* name, cudaStream_t s) {
  if (!g_on) return;
  g_cur_name = name;
  g_cur_a = get_event();
  cudaEventRecord(g_cur_a, s);
}
void ttl_prof_end(cudaStream_t s) {
  if (!g_on || !g_cur_a) return;
  cudaEvent_t b = get_event();
  cudaEventRecord(b, s);
  g_pending.push_back({g_cur_name, g_cur_a, b});
  g_cur_a = nullptr;
}

extern "C" {

void ttl_prof_enable(int32_t on) { g_on = on != 0; }
void ttl_pdl_enable(int32_t on) { g_ttl_pdl.store(on != 0); }

// Synchronises the recorded events, writes {"kernel": [launches, total_ms], ...} as JSON into
// buf (NUL-terminated, truncated to buflen) and clears the record.  Returns bytes needed.
int32_t ttl_prof_report(char* buf, int32_t buflen) {
  std::map<std::string, std::pair<long long, double>> acc;
  for (auto& p : g_pending) {
    cudaEventSynchronize(p.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.a, p.b);
    auto& e = acc[p.name];
    e.first += 1;
    e.second += ms;
    g_pool.push_back(p.a);
    g_pool.push_back(p.b);
  }
  g_pending.clear();
  std::string out = "{";
  bool first = true;
  for (auto& kv : acc) {
    char tmp[256];
    snprintf(tmp, sizeof(tmp), "%s\"%s\": [%lld, %.6f]", first ? "" : ", ", kv.first.c_str(),
             kv.second.first, kv.second.second);
    out += tmp;
    first = false;
  }
  out += "}";
  if (buf && buflen > 0) {
    const size_t n = out.size() < (size_t)buflen - 1 ? out.size() : (size_t)buflen - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int32_t)out.size() + 1;
}

}  // extern "C"
