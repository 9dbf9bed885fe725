// model.cpp -- see model.hpp.  Host only (g++).
#include "model.hpp"

#include <cctype>
#include <cstdlib>
#include <cstring>

#ifndef NNB_WEIGHTS_PATH
#error "NNB_WEIGHTS_PATH must point at data/weights.rnn"
#endif

// Embed the built-in model (the reference does the same with include_bytes!, src/rnn.rs:237).
__asm__(".section .rodata\n"
        ".global nnb_builtin_weights\n"
        ".global nnb_builtin_weights_end\n"
        ".balign 16\n"
        "nnb_builtin_weights:\n"
        ".incbin \"" NNB_WEIGHTS_PATH "\"\n"
        "nnb_builtin_weights_end:\n"
        ".byte 0\n"
        ".text\n");
extern "C" const unsigned char nnb_builtin_weights[];
extern "C" const unsigned char nnb_builtin_weights_end[];

namespace nnb {

namespace {
struct Cursor {
    const int8_t* base;
    size_t pos, len;
    size_t left() const { return len - pos; }
};

bool read_header(Cursor& c, int* ni, int* nn, int* act) {
    if (c.left() < 3) return false;
    const int8_t* b = c.base + c.pos;
    // header byte order is [nb_inputs, nb_neurons, activation] (src/rnn.rs:150-152, 171-173);
    // negative sizes are rejected by `unsigned` (src/rnn.rs:128-134)
    if (b[0] < 0 || b[1] < 0) return false;
    if (b[2] < 0 || b[2] > 2) return false;  // src/rnn.rs:136-143
    *ni = b[0];
    *nn = b[1];
    *act = b[2];
    c.pos += 3;
    return true;
}

bool take(Cursor& c, size_t n, size_t* off) {
    if (c.left() < n) return false;
    *off = c.pos;
    c.pos += n;
    return true;
}

bool read_dense(Cursor& c, HostDense* l) {
    return read_header(c, &l->ni, &l->nn, &l->act) && take(c, (size_t)l->ni * l->nn, &l->w_off) &&
           take(c, (size_t)l->nn, &l->b_off);
}

bool read_gru(Cursor& c, HostGru* l) {
    return read_header(c, &l->ni, &l->nn, &l->act) && take(c, (size_t)3 * l->nn * l->ni, &l->w_off) &&
           take(c, (size_t)3 * l->nn * l->nn, &l->r_off) && take(c, (size_t)3 * l->nn, &l->b_off);
}
}  // namespace

bool HostModel::parse(const uint8_t* data, size_t len, HostModel* m) {
    m->bytes.assign(reinterpret_cast<const int8_t*>(data), reinterpret_cast<const int8_t*>(data) + len);
    Cursor c{m->bytes.data(), 0, len};
    // layer order: src/rnn.rs:189-194
    bool ok = read_dense(c, &m->input_dense) && read_gru(c, &m->vad_gru) && read_gru(c, &m->noise_gru) &&
              read_gru(c, &m->denoise_gru) && read_dense(c, &m->denoise_output) && read_dense(c, &m->vad_output);
    if (!ok || c.left() != 0) return false;  // src/rnn.rs:196-198
    // src/rnn.rs:204-222
    if (m->input_dense.ni != 42 || m->denoise_output.nn != 22 || m->vad_output.nn != 1) return false;
    if (m->input_dense.nn != m->vad_gru.ni || m->vad_gru.nn != m->vad_output.ni) return false;
    if (42 + m->input_dense.nn + m->vad_gru.nn != m->noise_gru.ni) return false;
    if (42 + m->vad_gru.nn + m->noise_gru.nn != m->denoise_gru.ni) return false;
    if (m->denoise_gru.nn != m->denoise_output.ni) return false;
    return true;
}

bool HostModel::parse_text(const char* text, size_t len, HostModel* out) {
    static const char kHeader[] = "rnnoise-nu model file version 1";
    size_t eol = 0;
    while (eol < len && text[eol] != '\n') eol++;
    size_t h = eol;
    while (h > 0 && std::isspace((unsigned char)text[h - 1])) h--;
    size_t b = 0;
    while (b < h && std::isspace((unsigned char)text[b])) b++;
    if (h - b != sizeof(kHeader) - 1 || std::memcmp(text + b, kHeader, h - b) != 0) return false;
    std::vector<uint8_t> bin;
    size_t i = eol;
    while (i < len) {
        while (i < len && std::isspace((unsigned char)text[i])) i++;
        if (i >= len) break;
        bool neg = false;
        if (text[i] == '-' || text[i] == '+') {
            neg = text[i] == '-';
            i++;
        }
        if (i >= len || !std::isdigit((unsigned char)text[i])) return false;
        long v = 0;
        while (i < len && std::isdigit((unsigned char)text[i])) {
            v = v * 10 + (text[i] - '0');
            if (v > 1000000) return false;
            i++;
        }
        if (neg) v = -v;
        long mod = ((v % 256) + 256) % 256;  // python's int(s) % 256
        bin.push_back((uint8_t)mod);
    }
    return parse(bin.data(), bin.size(), out);
}

const HostModel& HostModel::builtin() {
    static const HostModel* m = [] {
        HostModel* p = new HostModel();
        bool ok = HostModel::parse(nnb_builtin_weights, (size_t)(nnb_builtin_weights_end - nnb_builtin_weights), p);
        if (!ok) std::abort();  // the reference unwrap()s (src/rnn.rs:238)
        return p;
    }();
    return *m;
}

}  // namespace nnb
