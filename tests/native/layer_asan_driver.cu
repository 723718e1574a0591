// AddressSanitizer driver for the layer-epilogue kernels' tile functions (csrc/sc_layer.cu).
// The kernels are sequences of __host__ __device__ functions; the sc_hostcheck_* entry points run them thread by thread on host
// buffers.  Here every buffer is an exact-size heap allocation and the translation unit is built with -fsanitize=address, so any
// out-of-bounds read or write of the index arithmetic the GPU will execute aborts the run.  Built and run by
// tests/test_layer_asan.py:  nvcc -Xcompiler -fsanitize=address -I include -I neuraloperator_b200/csrc layer_asan_driver.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../neuraloperator_b200/csrc/sc_layer.cu"

// the three symbols sc_layer.cu takes from the rest of the library
namespace sc {
static std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
bool cuda_ok(cudaError_t e, const char*) { return e == cudaSuccess; }
std::atomic<uint64_t> g_launches{0};
}  // namespace sc

struct Buf {   // exact-size heap block (ASan red zones on both sides)
  float* p;
  explicit Buf(size_t n) : p(n ? static_cast<float*>(std::malloc(n * sizeof(float))) : nullptr) {
    for (size_t i = 0; i < n; ++i) p[i] = 0.25f * static_cast<float>((i * 2654435761u) % 17) - 2.0f;
  }
  ~Buf() { std::free(p); }
};

static int fails = 0;
#define CHECK(expr) do { if ((expr) != 0) { std::printf("FAILED: %s (%s)\n", #expr, sc::g_err.c_str()); ++fails; } } while (0)

int main(int argc, char** argv) {
  // full sweep (~9 min under ASan, 2847 cases; run once per change of sc_layer.cu) or, with "quick", the subset the test tier runs
  const bool quick = argc > 1 && std::string(argv[1]) == "quick";
  const std::vector<int> Bs = quick ? std::vector<int>{2} : std::vector<int>{1, 2, 3};
  const std::vector<int> Cs = quick ? std::vector<int>{1, 17, 65} : std::vector<int>{1, 3, 15, 16, 17, 63, 64, 65, 130};
  const std::vector<long long> Ps = quick ? std::vector<long long>{1, 33, 129, 2049, 4097}
                                          : std::vector<long long>{1, 31, 32, 33, 127, 128, 129, 2047, 2048, 2049, 4096, 4097};
  long long cases = 0;
  for (int B : Bs)
    for (int Ci : Cs)
      for (int Co : Cs)
        for (long long P : Ps) {
          if ((long long)B * (Ci + Co) * P > 1500000) continue;
          Buf in((size_t)B * Ci * P), w((size_t)Co * Ci), bias(Co), add((size_t)B * Co * P), gate(Co), gated((size_t)B * Co * P);
          Buf out((size_t)B * Co * P), pre((size_t)B * Co * P), din((size_t)B * Ci * P), dw((size_t)Co * Ci), db(Co), dg(Co), dgd((size_t)B * Co * P);
          CHECK(sc_hostcheck_channel_mix(in.p, w.p, Ci, 1, bias.p, add.p, gate.p, gated.p, SC_ACT_GELU, out.p, pre.p, B, Ci, Co, P));
          CHECK(sc_hostcheck_channel_mix(in.p, w.p, Ci, 1, nullptr, nullptr, nullptr, nullptr, SC_ACT_IDENTITY, out.p, nullptr, B, Ci, Co, P));
          CHECK(sc_hostcheck_channel_mix(nullptr, nullptr, 0, 0, bias.p, add.p, nullptr, gated.p, SC_ACT_GELU, out.p, pre.p, B, 0, Co, P));
          // input gradient: weight read through transposed strides, roles of Ci / Co swapped
          CHECK(sc_hostcheck_channel_mix(out.p, w.p, 1, Ci, nullptr, nullptr, nullptr, nullptr, SC_ACT_IDENTITY, din.p, nullptr, B, Co, Ci, P));
          CHECK(sc_hostcheck_channel_mix_act_backward(out.p, pre.p, SC_ACT_GELU, gate.p, gated.p, add.p, dgd.p, db.p, dg.p, B, Co, P));
          CHECK(sc_hostcheck_channel_mix_act_backward(out.p, nullptr, SC_ACT_IDENTITY, nullptr, nullptr, nullptr, nullptr, db.p, nullptr, B, Co, P));
          CHECK(sc_hostcheck_channel_mix_weight_grad(add.p, in.p, dw.p, B, Ci, Co, P));
          CHECK(sc_hostcheck_pointwise(SC_POINTWISE_TANH, out.p, nullptr, pre.p, (long long)B * Co * P));
          CHECK(sc_hostcheck_pointwise(SC_POINTWISE_TANH_BACKWARD, out.p, pre.p, out.p, (long long)B * Co * P));
          CHECK(sc_hostcheck_pointwise(SC_POINTWISE_ROUND_HALF, out.p, nullptr, out.p, (long long)B * Co * P));
          CHECK(sc_hostcheck_pointwise(SC_POINTWISE_MUL, out.p, pre.p, out.p, (long long)B * Co * P));
          if ((((long long)B * Co * P) & 1) == 0) {      // the complex-pair ops read the neighbour inside a pair
            CHECK(sc_hostcheck_pointwise(SC_POINTWISE_ADD_I_TIMES, out.p, pre.p, add.p, (long long)B * Co * P));
            CHECK(sc_hostcheck_pointwise(SC_POINTWISE_MUL_NEG_I, out.p, nullptr, add.p, (long long)B * Co * P));
          }
          ++cases;
        }
  std::printf("layer_asan_driver: %lld shape cases, %d failures\n", cases, fails);
  return fails ? 1 : 0;
}
