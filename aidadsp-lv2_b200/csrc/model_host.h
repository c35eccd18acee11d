// model_host.h -- model files -> packed weight blocks.
//
// Accepts the reference's JSON schema (exporter field names from
// modules/RTNeural/python/model_utils.py:66-101) with the acceptance rules of
// RtNeuralGeneric::loadModelFromPath (rt-neural-generic.cpp:963-1085), custom_model_creator
// (model_variant.hpp:62-71, 656-875) and the RTNeural weight loaders (RTN/model_loader.h:21-43,
// 149-196, 227-263): weight rows larger than the declared shape are an error (the reference's
// vector::at throws, which is what modules/RTNeural/models/bad_lstm.json exercises), shorter
// rows are zero-filled.
#pragma once

#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "dsp_host.h"
#include "json_min.h"

namespace aidax {

struct HostModel {
    aidax_model_desc desc{};
    std::vector<float> W;  // [in][G*H]
    std::vector<float> U;  // [H][G*H]
    std::vector<float> b;  // LSTM [4H]; GRU [2][3H]
    std::vector<float> dense_w; // [H]
    float dense_b = 0.f;
    std::string path;
    // the exporter's self-test vectors, when the file holds them (testModel, rt-neural-generic.cpp:900-955, 1066-1073)
    std::vector<float> test_in, test_out;
};

inline bool hidden_size_supported(int h)
{
    // hidden_sizes of variant/generate_variant_hpp.py:6
    static const int sizes[] = {8, 12, 16, 20, 24, 32, 40, 64, 80};
    for (int s : sizes) if (s == h) return true;
    return false;
}

inline void fill_matrix(const JsonValue& rows, int n_rows, int n_cols, std::vector<float>& out, const char* what)
{
    out.assign((size_t)n_rows * n_cols, 0.f);
    if (!rows.is_array()) throw std::runtime_error(std::string(what) + ": not an array");
    for (size_t i = 0; i < rows.size(); ++i) {
        const JsonValue& row = rows.at(i);
        if (!row.is_array()) throw std::runtime_error(std::string(what) + ": row is not an array");
        for (size_t j = 0; j < row.size(); ++j) {
            if ((int)i >= n_rows || (int)j >= n_cols) throw std::out_of_range(std::string(what) + ": more weights than the declared shape");
            out[i * n_cols + j] = row.at(j).as_float();
        }
    }
}

// Throws std::exception on anything the reference would refuse.
inline HostModel parse_model_json(const std::string& text)
{
    const JsonValue j = json_parse(text);
    HostModel m;

    const int input_size = j["in_shape"].back().as_int();
    if (input_size > kMaxInput) throw std::invalid_argument("Value for input_size not supported");
    int in_skip = 0;
    if (j["in_skip"].is_number()) {
        in_skip = j["in_skip"].as_int();
        if (in_skip > 1) throw std::invalid_argument("Values for in_skip > 1 are not supported");
    }
    const float in_gain = j["in_gain"].is_number() ? db_to_coef(j["in_gain"].as_float()) : 1.0f;
    const float out_gain = j["out_gain"].is_number() ? db_to_coef(j["out_gain"].as_float()) : 1.0f;
    float samplerate = 48000.0f;
    if (j["metadata"]["samplerate"].is_number()) samplerate = j["metadata"]["samplerate"].as_float();
    else if (j["samplerate"].is_number()) samplerate = j["samplerate"].as_float();

    const JsonValue& layers = j.at("layers");
    const JsonValue& rnn = layers.at((size_t)0);
    const std::string& type = rnn.at("type").as_string();
    const int hidden = rnn.at("shape").back().as_int();
    int cell;
    if (type == "lstm") cell = AIDAX_CELL_LSTM;
    else if (type == "gru") cell = AIDAX_CELL_GRU;
    else throw std::runtime_error("Unable to identify a known model architecture!");
    if (!hidden_size_supported(hidden) || input_size < 1)
        throw std::runtime_error("Unable to identify a known model architecture!");

    m.desc.cell = cell;
    m.desc.hidden = hidden;
    m.desc.input_size = input_size;
    m.desc.in_skip = in_skip != 0;
    m.desc.in_gain = in_gain;
    m.desc.out_gain = out_gain;
    m.desc.samplerate = samplerate;

    const int G = cell == AIDAX_CELL_LSTM ? 4 : 3;
    const JsonValue& w = rnn["weights"];
    fill_matrix(w.at((size_t)0), input_size, G * hidden, m.W, "kernel weights");
    fill_matrix(w.at((size_t)1), hidden, G * hidden, m.U, "recurrent weights");
    if (cell == AIDAX_CELL_LSTM) {
        const JsonValue& bias = w.at((size_t)2);
        if (!bias.is_array() || (int)bias.size() < 4 * hidden) throw std::out_of_range("lstm bias shorter than 4*hidden");
        m.b.resize((size_t)4 * hidden);
        for (int k = 0; k < 4 * hidden; ++k) m.b[k] = bias.at((size_t)k).as_float();
    } else {
        fill_matrix(w.at((size_t)2), 2, 3 * hidden, m.b, "gru bias");
    }

    // ModelT's second layer is DenseT<float,H,1>: type must be dense and out dim 1, otherwise the
    // reference silently keeps zero weights (checkDense fails, model_loader.h:55-71).  A silently
    // silent amp is never what the user wants, so this is an error here.
    const JsonValue& dl = layers.at((size_t)1);
    const std::string& dtype = dl.at("type").as_string();
    if (dtype != "dense" && dtype != "time-distributed-dense") throw std::runtime_error("second layer is not dense");
    if (dl.at("shape").back().as_int() != 1) throw std::runtime_error("dense layer output size is not 1");
    const JsonValue& dw = dl["weights"];
    // loadDense: weights[0] is [in][out]; element (i, j) -> denseWeights[j][i]
    std::vector<float> tmp;
    fill_matrix(dw.at((size_t)0), hidden, 1, tmp, "dense weights");
    m.dense_w = tmp;
    const JsonValue& db = dw.at((size_t)1);
    if (!db.is_array() || db.size() < 1) throw std::out_of_range("dense bias missing");
    m.dense_b = db.at((size_t)0).as_float();

    if (j["input_batch"].is_array() && j["output_batch"].is_array()) {
        const JsonValue& xi = j["input_batch"];
        const JsonValue& yo = j["output_batch"];
        const size_t n = std::min(xi.size(), yo.size());
        m.test_in.resize(n);
        m.test_out.resize(n);
        for (size_t i = 0; i < n; ++i) {
            m.test_in[i] = xi.at(i).as_float();
            m.test_out[i] = yo.at(i).as_float();
        }
    }
    return m;
}

inline HostModel load_model_file(const char* path)
{
    std::ifstream f(path, std::ifstream::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    std::stringstream ss;
    ss << f.rdbuf();
    HostModel m = parse_model_json(ss.str());
    m.path = path;
    return m;
}

// Pack into the kernel layout documented in layout.h.  Gate order inside the JSON 4H/3H axis is
// Keras order: LSTM i,f,c,o (lstm_xsimd.tpp:216-219), GRU z,r,h (gru_xsimd.tpp:295-297); the two
// GRU bias rows are summed for z and r and kept apart for h (gru_xsimd.tpp:330-333).
inline std::vector<float> pack_model(const HostModel& m)
{
    const int H = m.desc.hidden, IN = m.desc.input_size;
    const bool lstm = m.desc.cell == AIDAX_CELL_LSTM;
    const int G = lstm ? 4 : 3;
    std::vector<float> p((size_t)packed_weight_floats(H), 0.f);
    float* wk = p.data();
    float* bias = wk + (size_t)(H + 3) * H * 4;
    float* dw = bias + (size_t)H * 4;
    float* db = dw + H;
    for (int k = 0; k < H; ++k)
        for (int j = 0; j < H; ++j) {
            float* d = wk + ((size_t)k * H + j) * 4;
            const float* u = &m.U[(size_t)k * G * H];
            if (lstm) { d[0] = u[j]; d[1] = u[j + H]; d[2] = u[j + 2 * H]; d[3] = u[j + 3 * H]; }
            else { d[0] = u[j]; d[1] = u[j + H]; d[2] = u[j + 2 * H]; d[3] = 0.f; }
        }
    for (int i = 0; i < IN; ++i)
        for (int j = 0; j < H; ++j) {
            float* d = wk + ((size_t)(H + i) * H + j) * 4;
            const float* w = &m.W[(size_t)i * G * H];
            if (lstm) { d[0] = w[j]; d[1] = w[j + H]; d[2] = w[j + 2 * H]; d[3] = w[j + 3 * H]; }
            else { d[0] = w[j]; d[1] = w[j + H]; d[2] = 0.f; d[3] = w[j + 2 * H]; }
        }
    for (int j = 0; j < H; ++j) {
        float* d = bias + (size_t)j * 4;
        if (lstm) { d[0] = m.b[j]; d[1] = m.b[j + H]; d[2] = m.b[j + 2 * H]; d[3] = m.b[j + 3 * H]; }
        else {
            const float* b0 = &m.b[0];
            const float* b1 = &m.b[(size_t)3 * H];
            d[0] = b0[j] + b1[j];
            d[1] = b0[j + H] + b1[j + H];
            d[2] = b1[j + 2 * H]; // bh1, added to U_h.h before the reset gate multiplies
            d[3] = b0[j + 2 * H]; // bh0, added to W_h.x
        }
        dw[j] = m.dense_w[j];
    }
    db[0] = m.dense_b;
    return p;
}

} // namespace aidax
