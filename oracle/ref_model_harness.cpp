/*
 * TEST INFRASTRUCTURE (oracle tier B) -- not product code.
 *
 * Thin C-ABI wrapper around the UNMODIFIED reference model path so Python tests can call it
 * through ctypes: the generated variant table (rt-neural-generic/src/model_variant.hpp:6-60,
 * 656-875) + RTNeural ModelT::parseJson/reset/forward (RTN/ModelT.h:258-379).  It is compiled
 * by oracle/build_ref.sh from the sources where they lie under /root/reference into
 * oracle/_ref/ (git-ignored).  Nothing from the reference is copied into this repository.
 *
 * The per-sample loop mirrors RtNeuralGeneric::applyModel (rt-neural-generic.cpp:148-240)
 * minus the gains/skip (those are applied by the caller or left at identity) so the raw cell
 * is observable; `padded` inputs are 4 floats per sample so the xsimd aligned full-register
 * load of ModelT::forward (ModelT.h:263-264) never sees unspecified lanes.
 */
#include <fstream>
#include <iostream>
#include <sstream>
#include <cstring>
#include <model_variant.hpp>

struct RefModel {
    ModelVariantType variant;
    int input_size = 0;
};

extern "C" {

void* ref_model_create(const char* json_path, int* input_size_out)
{
    try {
        nlohmann::json j;
        std::ifstream s(json_path, std::ifstream::binary);
        s >> j;
        auto* m = new RefModel();
        if (!custom_model_creator(j, m->variant)) { delete m; return nullptr; }
        // parseJson(debug=true) prints to std::cout in the reference; silence it here
        std::streambuf* old = std::cout.rdbuf();
        std::ostringstream sink;
        std::cout.rdbuf(sink.rdbuf());
        std::visit([&j](auto&& mdl) {
            using M = std::decay_t<decltype(mdl)>;
            if constexpr (!std::is_same_v<M, NullModel>) { mdl.parseJson(j, false); mdl.reset(); }
        }, m->variant);
        std::cout.rdbuf(old);
        m->input_size = j["in_shape"].back().get<int>();
        if (input_size_out) *input_size_out = m->input_size;
        return m;
    } catch (const std::exception& e) {
        std::cerr << "ref_model_create: " << e.what() << std::endl;
        return nullptr;
    }
}

void ref_model_reset(void* h)
{
    auto* m = (RefModel*)h;
    std::visit([](auto&& mdl) {
        using M = std::decay_t<decltype(mdl)>;
        if constexpr (!std::is_same_v<M, NullModel>) mdl.reset();
    }, m->variant);
}

/* in: [n][4] floats (lane 0 audio, 1 param1, 2 param2, 3 zero pad); out: [n] = forward(in). */
void ref_model_forward(void* h, const float* in_padded, float* out, int n)
{
    auto* m = (RefModel*)h;
    std::visit([&](auto&& mdl) {
        using M = std::decay_t<decltype(mdl)>;
        if constexpr (!std::is_same_v<M, NullModel>) {
            alignas(32) float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < n; ++i) {
                std::memcpy(v, in_padded + 4 * (size_t)i, 4 * sizeof(float));
                out[i] = mdl.forward(v);
            }
        }
    }, m->variant);
}

void ref_model_destroy(void* h) { delete (RefModel*)h; }

const char* ref_model_flavour()
{
#if RTNEURAL_USE_XSIMD
#  if defined(__AVX2__)
    return "xsimd-avx2";
#  else
    return "xsimd-sse2";
#  endif
#elif RTNEURAL_USE_EIGEN
    return "eigen";
#else
    return "stl";
#endif
}

} // extern "C"
