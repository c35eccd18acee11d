// aidax-batch-host: command-line batch renderer.  Instantiates N instances of an rt-neural-generic
// LV2 binary (this repository's CUDA-backed plugin by default, or the unmodified reference build),
// loads one model JSON into all of them through the patch:Set atom message, renders a synthetic
// sweep for the requested duration in buffer-sized run() calls and reports the real-time margin.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/aidax_host.h"

static void usage()
{
    fprintf(stderr,
            "usage: aidax-batch-host --plugin <rt-neural-generic.so> --model <file.json> [--instances N] [--seconds S]\n"
            "                        [--buffer 128] [--threads T] [--device D] [--eq-pre] [--dump out.f32]\n");
}

int main(int argc, char** argv)
{
    std::string plugin, model, dump;
    int n = 1024, buffer = 128, threads = 1, device = -1;
    double seconds = 1.0;
    bool eq_pre = false;
    for (int i = 1; i < argc; ++i) {
        auto need = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { fprintf(stderr, "%s needs a value\n", name); exit(2); }
            return argv[++i];
        };
        if (!strcmp(argv[i], "--plugin")) plugin = need("--plugin");
        else if (!strcmp(argv[i], "--model")) model = need("--model");
        else if (!strcmp(argv[i], "--instances")) n = atoi(need("--instances"));
        else if (!strcmp(argv[i], "--seconds")) seconds = atof(need("--seconds"));
        else if (!strcmp(argv[i], "--buffer")) buffer = atoi(need("--buffer"));
        else if (!strcmp(argv[i], "--threads")) threads = atoi(need("--threads"));
        else if (!strcmp(argv[i], "--device")) device = atoi(need("--device"));
        else if (!strcmp(argv[i], "--dump")) dump = need("--dump");
        else if (!strcmp(argv[i], "--eq-pre")) eq_pre = true;
        else { usage(); return 2; }
    }
    if (plugin.empty() || model.empty()) { usage(); return 2; }
    const double fs = 48000.0;
    aidax_host* h = aidax_host_open(plugin.c_str(), fs, n, buffer, threads, device);
    if (!h) { fprintf(stderr, "cannot open host: %s\n", aidax_host_error()); return 1; }
    if (eq_pre) aidax_host_set_control(h, -1, 10, 1.0f);
    aidax_host_load_model(h, -1, model.c_str());
    aidax_host_activate(h);
    float* in = aidax_host_input(h);
    float* out = aidax_host_output(h);
    const size_t stride = aidax_host_stride(h);
    const long total = (long)(seconds * fs);
    FILE* df = dump.empty() ? nullptr : fopen(dump.c_str(), "wb");
    double busy = 0.0;
    long done = 0;
    for (long t0 = 0; t0 < total; t0 += buffer) {
        const int nb = (int)std::min<long>(buffer, total - t0);
        for (int i = 0; i < n; ++i) {
            float* row = in + (size_t)i * stride;
            for (int k = 0; k < nb; ++k) {
                const double t = (double)(t0 + k) / fs;
                row[k] = (float)(0.5 * std::sin(2.0 * M_PI * (20.0 * t + (20000.0 - 20.0) / 20.0 * t * t) + 0.37 * i));
            }
        }
        const auto a = std::chrono::steady_clock::now();
        if (aidax_host_run(h, nb) != 0) { fprintf(stderr, "run failed: %s\n", aidax_host_error()); return 1; }
        busy += std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
        if (df) fwrite(out, sizeof(float), (size_t)nb, df); // instance 0
        done += nb;
    }
    if (df) fclose(df);
    const double audio = (double)done / fs;
    printf("{\"plugin\": \"%s\", \"batched\": %d, \"instances\": %d, \"buffer\": %d, \"threads\": %d, \"audio_s\": %.3f, "
           "\"busy_s\": %.4f, \"us_per_buffer\": %.2f, \"realtime_factor\": %.2f, \"model_in_size\": %.0f, \"notifications\": %d, "
           "\"gpu_launches\": %llu}\n",
           plugin.c_str(), aidax_host_is_batched(h), n, buffer, threads, audio, busy, busy / (double)((done + buffer - 1) / buffer) * 1e6,
           audio / busy, aidax_host_get_control(h, 0, 23), aidax_host_notify_count(h, 0), (unsigned long long)aidax_host_launch_count(h));
    aidax_host_close(h);
    return 0;
}
