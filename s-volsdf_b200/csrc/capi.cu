// Library-wide C-ABI helpers: last-error string, version, engine query, launch counter and the optional
// per-kernel profiler (CUDA events on the launching stream, aggregated by kernel name).
#include <stdarg.h>

#include <map>
#include <string>
#include <vector>

#include "svs_common.cuh"

namespace svs {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static long long g_launches = 0;
void count_launch() { ++g_launches; }

struct ProfRec {
  const char* name;
  double flops, bytes;
  cudaEvent_t a, b;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;

bool prof_enabled() { return g_prof_on; }

int prof_begin(const char* name, double flops, double bytes, cudaStream_t st) {
  if (!g_prof_on) return -1;
  ProfRec r;
  r.name = name;
  r.flops = flops;
  r.bytes = bytes;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
  return (int)g_prof.size() - 1;
}

void prof_end(int id, cudaStream_t st) {
  if (id < 0 || id >= (int)g_prof.size()) return;
  cudaEventRecord(g_prof[id].b, st);
}
}  // namespace svs

extern "C" const char* svs_last_error(void) { return svs::g_err; }
extern "C" int svs_abi_version(void) { return SVS_ABI_VERSION; }
extern "C" int svs_has_engine(int engine) {
  return engine == SVS_ENGINE_FP32 || engine == SVS_ENGINE_TC || engine == SVS_ENGINE_TC_SPLIT;
}

extern "C" int64_t svs_launch_count(void) { return svs::g_launches; }

extern "C" int svs_prof_enable(int on) {
  svs::g_prof_on = on != 0;
  return SVS_OK;
}

// Synchronises the recorded events, aggregates them by kernel name and writes one line per kernel:
//   name <tab> launches <tab> total_ms <tab> flops_per_launch_sum <tab> bytes_per_launch_sum
// Returns the number of bytes written (excluding the terminator) or <0 on error; clears the records.
extern "C" int64_t svs_prof_collect(char* buf, int64_t cap) {
  struct Agg {
    long long n = 0;
    double ms = 0, flops = 0, bytes = 0;
  };
  std::map<std::string, Agg> agg;
  for (auto& r : svs::g_prof) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      Agg& a = agg[r.name];
      a.n += 1;
      a.ms += ms;
      a.flops += r.flops;
      a.bytes += r.bytes;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  svs::g_prof.clear();
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s\t%lld\t%.6f\t%.6e\t%.6e\n", kv.first.c_str(), kv.second.n, kv.second.ms,
             kv.second.flops, kv.second.bytes);
    out += line;
  }
  if ((int64_t)out.size() + 1 > cap) {
    svs::set_error("svs_prof_collect: buffer too small (%lld needed)", (long long)out.size() + 1);
    return SVS_ERR_INVALID;
  }
  memcpy(buf, out.c_str(), out.size() + 1);
  return (int64_t)out.size();
}
