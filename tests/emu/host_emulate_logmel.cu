// Host emulation of logmel_kernel (CPU, no GPU needed): runs the exact __host__ __device__ phase functions of
// csrc/logmel.cuh thread by thread, in the kernel's order, so the index math is tested in the no-GPU container.
// TEST INFRASTRUCTURE ONLY — never linked into libwhisper_b200.so.
#include "../../openai-whisper-coreml_b200/csrc/logmel.cu"

#include <vector>

namespace {
template <typename T>
void emulate(const T* clip /* unpadded [480000] */, T* out /* [80][3000] */) {
  using Cfg = wb::LogmelCfg<T>;
  constexpr int F = Cfg::F, NT = Cfg::NT;
  auto* sm = new wb::LogmelSmem<T, F>();
  wb::build_logmel_tables<T>(sm->tab);
  std::vector<T> logspec(80 * 3000);
  T gmax = (T)-1e30;
  for (int tile = 0; tile < 3000 / F; ++tile) {
    const int p0 = tile * F * WB_HOP;
    for (int s = 0; s < wb::tile_samples<F>(); ++s) sm->region0[wb::samp_index(s)] = clip[wb::reflect_index(p0 + s)];
    for (int tid = 0; tid < NT; ++tid) wb::logmel_phase_a<T, F>(*sm, tid);
    for (int task = 0; task < F * 25; ++task) wb::logmel_phase_b<T, F>(*sm, task);
    // the kernel's thread -> work maps of phases C1 and C2 (a thread keeps its bin pair / its mel band and walks the frames)
    constexpr int FS = NT / 100;
    for (int tid = 0; tid < 100 * FS; ++tid)
      for (int fl = tid / 100; fl < F; fl += FS) wb::logmel_phase_c1<T, F>(*sm, fl, tid % 100 + 1);
    constexpr int GF = NT / 80, NF = (F + GF - 1) / GF;
    for (int tid = 0; tid < 80 * GF; ++tid) {
      const int i = tid / GF, g = tid - i * GF;
      T v[NF];
      wb::logmel_phase_c2<T, F, NF>(*sm, i, g, GF, v);
      for (int n = 0; n < NF; ++n) {
        const int fl = g + GF * n;
        if (fl >= F) continue;
        logspec[i * 3000 + tile * F + fl] = v[n];
        gmax = v[n] > gmax ? v[n] : gmax;
      }
    }
  }
  for (int i = 0; i < 80 * 3000; ++i) {
    T v = logspec[i];
    const T fl = gmax - (T)8.0;
    v = v > fl ? v : fl;
    out[i] = (v + (T)4.0) / (T)4.0;
  }
  delete sm;
}
}  // namespace

extern "C" void emu_logmel_f32(const float* clip, float* out) { emulate<float>(clip, out); }
extern "C" void emu_logmel_f64(const double* clip, double* out) { emulate<double>(clip, out); }
// stand-ins for the library's error plumbing (csrc/api.cu) so this test object links on its own
namespace wb {
void set_error(const char*, ...) {}
const char* get_error() { return ""; }
}  // namespace wb
