// BMCNet / BMCNet_plain forward as a static plan of kernel launches over a caller-provided arena.
//
// Reference: models/BMCNet.py:19-121, models/BMCNet_plain.py:3-68, models/submodules.py:17-77.
// The module tree (and therefore the state_dict) lives on the Python side; this file knows the
// dataflow.  A plan is built once per (model, batch, H, W): every activation is a 128-channel
// bf16 "slot" of the padded NHWC arena, every convolution / BIE product is one conv-gemm job
// (gemm.cuh), independent jobs of equal shape share a launch (grid.y), and the whole step is
// replayed as one CUDA graph.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "gemm.cuh"

using namespace bmc;

namespace {

// ------------------------------------------------------------------ state_dict bookkeeping
std::vector<std::string> conv_keys(const std::string& p) { return {p + ".weight", p + ".bias"}; }
void append(std::vector<std::string>& a, const std::vector<std::string>& b) { a.insert(a.end(), b.begin(), b.end()); }
std::vector<std::string> res_keys(const std::string& p) {
    auto k = conv_keys(p + ".conv1");
    append(k, conv_keys(p + ".conv2"));
    return k;
}
std::vector<std::string> bie_keys(const std::string& p) {
    auto k = res_keys(p + ".conv1");
    append(k, res_keys(p + ".conv2"));
    append(k, conv_keys(p + ".convf1"));
    append(k, conv_keys(p + ".convf2"));
    k.push_back(p + ".norm_s.weight");
    k.push_back(p + ".norm_s.bias");
    for (const char* n : {"clustering", "unclustering", "v1", "v2"}) append(k, conv_keys(p + "." + n));
    return k;
}
// The reference key set in module order: 318 keys for BMCNet, 120 for plain.
std::vector<std::string> canonical_keys(int kind, int n_b) {
    std::vector<std::string> k;
    if (kind == BMC_MODEL_BMCNET_PLAIN) {
        for (const char* n : {"conv_f1", "conv_f2", "conv_fs"}) append(k, conv_keys(std::string("neuro.") + n));
        for (int i = 0; i < n_b; ++i) append(k, bie_keys("neuro.para_reschunk." + std::to_string(i)));
        for (const char* n : {"conv_h", "conv_o"}) append(k, conv_keys(std::string("neuro.") + n));
    } else {
        for (const char* n : {"conv_fpst", "conv_fnst", "conv_fps", "conv_fns", "conv_fs"})
            append(k, conv_keys(std::string("neuro.") + n));
        for (int i = 0; i < n_b; ++i) {
            const std::string p = "neuro.para_reschunk." + std::to_string(i);
            for (const char* n : {"conv1", "conv2", "conv1_st", "conv2_st"}) append(k, res_keys(p + "." + n));
            append(k, bie_keys(p + ".lBIE"));
            append(k, bie_keys(p + ".gBIE"));
        }
        for (const char* n : {"conv_hs", "conv_hp", "conv_hn", "conv_o"}) append(k, conv_keys(std::string("neuro.") + n));
    }
    return k;
}
std::vector<std::string> split(const std::string& s) {
    std::vector<std::string> out;
    size_t a = 0;
    while (true) {
        size_t b = s.find('.', a);
        out.push_back(s.substr(a, b == std::string::npos ? b : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    return out;
}
// Owner of an aliased key: `[Blk] * n_b` repeats ONE block object, `conv2 = conv1` etc. alias
// sub-modules (BMCNet.py:6-9,41,43,46; BMCNet_plain.py:8,11; submodules.py:43,45).
std::string alias_root(const std::string& key) {
    auto p = split(key);
    if (p.size() > 2 && p[1] == "para_reschunk") p[2] = "0";
    auto is_handle = [](const std::string& s) { return s == "conv1" || s == "conv2" || s == "conv1_st" || s == "conv2_st"; };
    for (size_t i = 0; i < p.size(); ++i) {
        const bool res_leaf = (p[i] == "conv1" || p[i] == "conv2") && i + 2 == p.size() && i > 0 && is_handle(p[i - 1]);
        if (res_leaf) continue;
        if (p[i] == "conv2") p[i] = "conv1";
        else if (p[i] == "conv2_st") p[i] = "conv1_st";
        else if (p[i] == "convf2") p[i] = "convf1";
        else if (p[i] == "conv_fnst") p[i] = "conv_fpst";
        else if (p[i] == "conv_fns") p[i] = "conv_fps";
        else if (p[i] == "conv_f2") p[i] = "conv_f1";
    }
    std::string out = p[0];
    for (size_t i = 1; i < p.size(); ++i) out += "." + p[i];
    return out;
}

// ------------------------------------------------------------------ weight specs
struct SegSpec {
    std::vector<int> src;    // per padded channel: input channel of the reference conv, or -1
};
struct WeightSpec {
    std::string conv;        // state_dict prefix, e.g. "neuro.conv_f1"
    std::string fold_ln;     // non-empty: prefix of the LayerNorm whose affine (gamma, beta) is folded into this 1x1 conv
    int cin, n_out, taps;
    std::vector<SegSpec> segs;
    int row_base = 0, k_chunks = 0;
    size_t bias_off = 0;     // in floats, inside the fp32 region
};
struct LnSpec { std::string prefix; size_t gamma_off = 0, beta_off = 0; };

SegSpec seg_range(int first, int count, int pad_to) {
    SegSpec s;
    s.src.assign(pad_to, -1);
    for (int i = 0; i < count; ++i) s.src[i] = first + i;
    return s;
}
// the 64-channel per-step input tensor (pointwise.cu: pack_inputs); `at` pairs are
// (first MI channel, first source channel, count)
SegSpec seg_mi(std::initializer_list<std::array<int, 3>> at) {
    SegSpec s;
    s.src.assign(64, -1);
    for (auto& a : at)
        for (int i = 0; i < a[2]; ++i) s.src[a[0] + i] = a[1] + i;
    return s;
}

struct Src { int kind; int slot; };   // kind 0: 128-channel arena slot, 1: the MI tensor
struct JobSpec {
    std::vector<Src> segs;
    int weight = -1;         // index into weights, or
    int dyn_pair = -1;       // dynamic (softmax) weights of this attention pair
    bool relu = false;
    int res_slot = -1;
    int out_slot = -1;
    bool out_f32 = false;    // conv_o: fp32 [rows][32] side buffer
    int hid32 = -1;          // hidden-state convs: also keep an fp32 copy [rows][128] (forward() output)
    int ln = -1;
    // centre-tap-only segments appended after `segs` (slab kernel only, gemm.cuh GemmJobDev::t1_*):
    int mix_slot = -1, mix_pair = -1;   // + M_pair[b] . x  with x in arena slot mix_slot (fused BIE, bie_fused.cu)
    int ident_slot = -1;                // + I . x : a residual add done by the tensor core
    bool accumulate = false;            // out_slot already holds the residual: add the result onto it (conv_slab2_tc reduce-add epilogue)
};

struct Op {
    enum Kind { kGemm, kAtt, kSoftmax, kBieFront, kFold } kind;
    GemmParams gp;
    AttParams ap;
    SoftmaxParams sp;
    BieFrontParams fp;
    FoldParams dp;
};

}  // namespace

struct bmc_model {
    int kind, scale, n_c, n_b, repeat;
    std::vector<WeightSpec> weights;
    std::vector<LnSpec> lns;
    std::map<std::string, int> widx;
    int w_rows_total = 0, ident_row = 0;
    size_t f32_floats = 0, kmap_ints = 0;
    // device weights
    act_t* w_dev = nullptr;
    float* f32_dev = nullptr;
    int* kmap_dev = nullptr;
    bool loaded = false;
    // geometry / workspace
    bool configured = false, bound = false;
    Geom g;
    int n_slots = 0, n_split = 1, pix_per_split = 0;
    size_t ws_bytes = 0;
    size_t off_mi = 0, off_a32 = 0, off_p = 0, off_partial = 0, off_hid32 = 0, off_bimg = 0, off_gpart = 0, off_spart = 0;
    bool fused = true;                     // product plan: fused BIE 1x1 / attention section (bie_fused.cu)
    bool allow_fused = true;
    bool inplace_res = true;               // ResidualBlock identity as an in-place reduce-add (needs conv_slab2_tc at this image size)
    char* ws = nullptr;
    CUtensorMap map_act, map_att, map_slab, map_mi, map_mi64, map_w128, map_w64, map_w32, map_p;
    CUtensorMap map_act_s64, map_mi_s64, map_w_s64, map_p_s64;   // 32-channel / 64-byte-swizzle boxes (conv_slab2_tc)
    CUtensorMap map_act_out;               // activation arena behind [32 rows x 64 ch] boxes (TMA-store epilogue)
    int abox = 64;                         // rows per TMA box of map_slab / map_mi64 (slab_box_rows)
    int abox32 = 64;                       // rows per TMA box of map_act_s64 / map_mi_s64 (slab2_box_rows)
    // plan
    std::vector<Op> ops;
    std::vector<int> free_slots;
    int slots_hi = 0;
    int slot_h[3] = {-1, -1, -1};
    bool dry = true;
    int simt = 0;
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraph_t graph = nullptr;
    cudaStream_t cap_stream = nullptr;     // capture happens here: the caller's stream may be the
                                           // legacy default stream, which cannot be captured
    int graph_simt = -1;
    int eager_runs = 0;
    bool use_graph = true;

    act_t* slot_ptr(int s) const { return reinterpret_cast<act_t*>(ws) + (size_t)s * g.rows() * 128; }
    act_t* mi_ptr() const { return reinterpret_cast<act_t*>(ws + off_mi); }
    float* a32_ptr() const { return reinterpret_cast<float*>(ws + off_a32); }
    act_t* p_ptr() const { return reinterpret_cast<act_t*>(ws + off_p); }
    float* partial_ptr() const { return reinterpret_cast<float*>(ws + off_partial); }
    float* hid32_ptr(int i) const { return reinterpret_cast<float*>(ws + off_hid32) + (size_t)i * g.rows() * 128; }
    float* bimg_ptr() const { return reinterpret_cast<float*>(ws + off_bimg); }
    float* gpart_ptr() const { return reinterpret_cast<float*>(ws + off_gpart); }
    float* spart_ptr() const { return reinterpret_cast<float*>(ws + off_spart); }
};

namespace {

// Aliased modules resolve to one entry: names are stored under their alias root.
std::string root_name(const std::string& name) {
    if (name.find('.') == std::string::npos) return name;
    const std::string r = alias_root(name + ".weight");
    return r.substr(0, r.size() - 7);
}

int add_weight(bmc_model* m, const std::string& name_in, const std::string& conv, int cin, int n_out, int taps,
               std::vector<SegSpec> segs) {
    const std::string name = root_name(name_in);
    if (m->widx.count(name)) return m->widx[name];
    WeightSpec w;
    w.conv = conv; w.cin = cin; w.n_out = n_out; w.taps = taps; w.segs = std::move(segs);
    int k = 0;
    for (auto& s : w.segs) k += (int)s.src.size() * taps;
    w.k_chunks = k / 64;
    w.row_base = m->w_rows_total;
    m->w_rows_total += w.k_chunks * n_out;
    w.bias_off = m->f32_floats;
    m->f32_floats += 128;
    m->kmap_ints += (size_t)k;
    m->widx[name] = (int)m->weights.size();
    m->weights.push_back(std::move(w));
    return (int)m->weights.size() - 1;
}

void add_res(bmc_model* m, const std::string& p) {
    add_weight(m, p + ".conv1", p + ".conv1", 128, 128, 9, {seg_range(0, 128, 128)});
    add_weight(m, p + ".conv2", p + ".conv2", 128, 128, 9, {seg_range(0, 128, 128)});
}
void add_bie(bmc_model* m, const std::string& p) {
    add_res(m, p + ".conv1");
    add_res(m, p + ".conv2");
    for (const char* n : {"convf1", "convf2", "unclustering"})
        add_weight(m, p + "." + n, p + "." + n, 256, 128, 1, {seg_range(0, 128, 128), seg_range(128, 128, 128)});
    for (const char* n : {"clustering", "v1", "v2"})
        add_weight(m, p + "." + n, p + "." + n, 128, 128, 1, {seg_range(0, 128, 128)});
    // clustering(norm_s(y)) = (Wc diag(gamma)) n + (bc + Wc beta) with n = (y - mu) * rstd: the fused front kernel
    // (bie_fused.cu) normalises without the affine part and multiplies by this folded copy
    {
        const int before = (int)m->weights.size();
        const int wi = add_weight(m, p + ".clustering_ln", p + ".clustering", 128, 128, 1, {seg_range(0, 128, 128)});
        if ((int)m->weights.size() > before) m->weights[wi].fold_ln = p + ".norm_s";
    }
    if (m->widx.count(root_name(p + ".norm_s"))) return;
    LnSpec ln;
    ln.prefix = p + ".norm_s";
    ln.gamma_off = m->f32_floats; m->f32_floats += 128;
    ln.beta_off = m->f32_floats; m->f32_floats += 128;
    m->widx[root_name(p + ".norm_s")] = (int)m->lns.size();
    m->lns.push_back(ln);
}

// Register every (weight, operand layout) pair the plan uses.  Channel maps follow the torch.cat
// orders of BMCNet.py:60-73 / BMCNet_plain.py:24-30 and the MI layout of pack_inputs.
void register_weights(bmc_model* m) {
    const std::string blk = "neuro.para_reschunk.";
    if (m->kind == BMC_MODEL_BMCNET_PLAIN) {
        // conv_f1(cat[in_1(6), h(128), o1(16)]) ; conv_f2 = same conv on (in_2, h, o2)
        add_weight(m, "f1", "neuro.conv_f1", 150, 128, 9, {seg_range(6, 128, 128), seg_mi({{0, 0, 6}, {12, 134, 16}})});
        add_weight(m, "f2", "neuro.conv_f2", 150, 128, 9, {seg_range(6, 128, 128), seg_mi({{6, 0, 6}, {28, 134, 16}})});
        // conv_fs(cat[in_1(6), in_2(6), h(128), o(32)])
        add_weight(m, "fs", "neuro.conv_fs", 172, 128, 9, {seg_range(12, 128, 128), seg_mi({{0, 0, 12}, {12, 140, 32}})});
        for (int i = 0; i < m->n_b; ++i) add_bie(m, blk + std::to_string(i));
        add_weight(m, "h", "neuro.conv_h", 128, 128, 9, {seg_range(0, 128, 128)});
        add_weight(m, "o", "neuro.conv_o", 256, 32, 9, {seg_range(0, 128, 128), seg_range(128, 128, 128)});
    } else {
        // conv_fpst(cat[x1p(3), x2p(3), hp(128), op(16)]) ; conv_fnst on the negative planes
        add_weight(m, "fpst", "neuro.conv_fpst", 150, 128, 9, {seg_range(6, 128, 128), seg_mi({{0, 0, 6}, {12, 134, 16}})});
        add_weight(m, "fnst", "neuro.conv_fnst", 150, 128, 9, {seg_range(6, 128, 128), seg_mi({{6, 0, 6}, {28, 134, 16}})});
        // conv_fps(cat[x2p(3), hp(128)])
        add_weight(m, "fps", "neuro.conv_fps", 131, 128, 9, {seg_range(3, 128, 128), seg_mi({{3, 0, 3}})});
        add_weight(m, "fns", "neuro.conv_fns", 131, 128, 9, {seg_range(3, 128, 128), seg_mi({{9, 0, 3}})});
        // conv_fs(cat[xp_st(128), xn_st(128), h*(128), o(32)])
        add_weight(m, "fs", "neuro.conv_fs", 416, 128, 9,
                   {seg_range(0, 128, 128), seg_range(128, 128, 128), seg_range(256, 128, 128), seg_mi({{12, 384, 32}})});
        for (int i = 0; i < m->n_b; ++i) {
            const std::string p = blk + std::to_string(i);
            for (const char* n : {"conv1", "conv2", "conv1_st", "conv2_st"}) add_res(m, p + "." + n);
            add_bie(m, p + ".lBIE");
            add_bie(m, p + ".gBIE");
        }
        for (const char* n : {"conv_hs", "conv_hp", "conv_hn"})
            add_weight(m, n, std::string("neuro.") + n, 128, 128, 9, {seg_range(0, 128, 128)});
        add_weight(m, "o", "neuro.conv_o", 256, 32, 9, {seg_range(0, 128, 128), seg_range(128, 128, 128)});
    }
    // 128x128 identity in the weight layout [2][128][64]: residual adds as one more K segment (gemm.cuh t1_*)
    m->ident_row = m->w_rows_total;
    m->w_rows_total += 256;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------ plan builder
struct Builder {
    bmc_model* m;
    int alloc() {
        if (!m->free_slots.empty()) { int s = m->free_slots.back(); m->free_slots.pop_back(); return s; }
        return m->slots_hi++;
    }
    void release(int s) { m->free_slots.push_back(s); }
    int W(const std::string& name) const { return m->widx.at(root_name(name)); }

    int gemm(const std::vector<JobSpec>& jobs, int n, int taps) {
        if (m->dry) return BMC_OK;
        Op op;
        op.kind = Op::kGemm;
        GemmParams& p = op.gp;
        memset(&p, 0, sizeof(p));
        const Geom& g = m->g;
        p.maps[0] = m->map_act; p.maps[1] = m->map_mi;
        p.maps[2] = (n == 32) ? m->map_w32 : m->map_w128;
        p.maps[3] = m->map_p;
        p.maps[4] = m->map_slab;                // act arena, slab boxes (slab / pair kernels)
        p.abox_rows = taps == 9 ? m->abox : 64;
        p.maps[5] = m->map_mi64;
        p.maps[6] = m->map_w64;
        p.maps[7] = m->map_att;                 // act arena, 64-row boxes (1x1 launches of the unfused plan)
        p.maps32[0] = m->map_act_s64; p.maps32[1] = m->map_mi_s64; p.maps32[2] = m->map_w_s64; p.maps32[3] = m->map_p_s64;
        p.maps32[4] = m->map_act_out;
        p.has32 = taps == 9; p.abox32_rows = m->abox32;
        p.n_jobs = (int)jobs.size();
        const int n_plain = (int)jobs[0].segs.size();
        p.n_seg = n_plain + (jobs[0].mix_slot >= 0) + (jobs[0].ident_slot >= 0);
        p.n_taps = taps;
        for (int t = 0; t < taps; ++t) p.tap_off[t] = taps == 9 ? (t / 3 - 1) * g.Wp + (t % 3 - 1) : 0;
        p.n = n; p.g = g; p.tiles_per_img = g.R / kTileM;
        for (int s = 0; s < p.n_seg; ++s) p.chunks[s] = (s < n_plain && jobs[0].segs[s].kind == 1) ? 1 : 2;
        for (int j = 0; j < p.n_jobs; ++j) {
            const JobSpec& js = jobs[j];
            GemmJobDev& d = p.jobs[j];
            std::vector<Src> segs = js.segs;
            if (js.mix_slot >= 0) {
                const int sg = (int)segs.size();
                segs.push_back({0, js.mix_slot});
                p.tap1_mask |= 1 << sg;
                d.t1_map[sg] = 3; d.t1_map32[sg] = 3; d.t1_row[sg] = js.mix_pair * g.B * 256; d.t1_img_stride[sg] = 256;
                d.bias_img = m->bimg_ptr() + (size_t)js.mix_pair * g.B * 128;
            }
            if (js.ident_slot >= 0) {
                const int sg = (int)segs.size();
                segs.push_back({0, js.ident_slot});
                p.tap1_mask |= 1 << sg;
                d.t1_map[sg] = 2; d.t1_map32[sg] = 2; d.t1_row[sg] = m->ident_row; d.t1_img_stride[sg] = 0;
            }
            for (int s = 0; s < p.n_seg; ++s) {
                const Src& src = segs[s];
                if (src.kind == 1) {
                    d.a_map[s] = 1; d.a_map64[s] = 5; d.a_map32[s] = 1; d.a_row_base[s] = 0; d.a_ptr[s] = m->mi_ptr(); d.a_ld[s] = 64; d.a_rows[s] = g.rows();
                } else {
                    d.a_map[s] = 0; d.a_map64[s] = taps == 9 ? 4 : 7; d.a_map32[s] = 0; d.a_row_base[s] = (int)(src.slot * g.rows());
                    d.a_ptr[s] = m->slot_ptr(0); d.a_ld[s] = 128; d.a_rows[s] = (long)m->n_slots * g.rows();
                }
                d.a_col_base[s] = 0;
            }
            if (js.weight >= 0) {
                const WeightSpec& w = m->weights[js.weight];
                d.w_map = 2; d.w_map64 = n == 128 ? 6 : 0; d.w_map32 = 2; d.w_ptr = m->w_dev; d.w_rows = w.n_out; d.w_row_base = w.row_base; d.w_img_stride = 0;
                d.bias = m->f32_dev + w.bias_off;
            } else {
                d.w_map = 3; d.w_map32 = -1; d.w_ptr = m->p_ptr(); d.w_rows = 128;
                d.w_row_base = js.dyn_pair * g.B * 256; d.w_img_stride = 256; d.bias = nullptr;
            }
            d.relu = js.relu ? 1 : 0;
            d.residual = js.res_slot >= 0 ? m->slot_ptr(js.res_slot) : nullptr;
            d.res_row_base = 0;
            d.out = js.out_slot >= 0 ? m->slot_ptr(js.out_slot) : nullptr;
            d.out_row_base = 0;
            d.out_map32 = js.out_slot >= 0 ? 4 : -1;
            d.out_map_row = js.out_slot >= 0 ? (int)(js.out_slot * g.rows()) : 0;
            d.out_f32 = js.out_f32 ? m->a32_ptr() : nullptr;
            d.out_accumulate = js.accumulate ? 1 : 0;
            if (js.ln >= 0) {
                d.ln_gamma = m->f32_dev + m->lns[js.ln].gamma_off;
                d.ln_beta = m->f32_dev + m->lns[js.ln].beta_off;
                d.ln_eps = 1e-6f;
            }
        }
        m->ops.push_back(op);
        return BMC_OK;
    }

    void attention(const std::vector<std::pair<int, int>>& cv) {   // (centres slot, v slot) per pair
        if (m->dry) return;
        const Geom& g = m->g;
        Op a;
        a.kind = Op::kAtt;
        memset(&a.ap, 0, sizeof(a.ap));
        a.ap.map_c = m->map_att; a.ap.map_v = m->map_att;
        a.ap.c_ptr = m->slot_ptr(0); a.ap.v_ptr = m->slot_ptr(0);
        a.ap.n_pairs = (int)cv.size(); a.ap.n_split = m->n_split; a.ap.pix_per_split = m->pix_per_split;
        a.ap.scale = 1.0f / sqrtf(128.f);                           // nf ** -0.5, submodules.py:47
        a.ap.partial = m->partial_ptr(); a.ap.g = g;
        for (size_t i = 0; i < cv.size(); ++i) {
            a.ap.c_row_base[i] = (long)cv[i].first * g.rows();
            a.ap.v_row_base[i] = (long)cv[i].second * g.rows();
        }
        m->ops.push_back(a);
        Op s;
        s.kind = Op::kSoftmax;
        memset(&s.sp, 0, sizeof(s.sp));
        s.sp.partial = m->partial_ptr(); s.sp.n_pairs = (int)cv.size(); s.sp.n_split = m->n_split; s.sp.B = g.B;
        s.sp.w_base = m->p_ptr(); s.sp.w_img_stride = 256;
        for (size_t i = 0; i < cv.size(); ++i) s.sp.w_row_base[i] = (int)i * g.B * 256;
        m->ops.push_back(s);
    }

    struct Tri { int x1, x2, xs; };
    // BIE.forward (submodules.py:58-77) for 1 or 2 independent instances sharing the weights.
    // Consumes (releases) its input slots.
    std::vector<Tri> bie(const std::string& p, const std::vector<Tri>& in) {
        if (m->fused) return bie_fused(p, in);
        const int I = (int)in.size();
        std::vector<std::array<int, 2>> t(I), r(I), y(I), c(I), v(I), nx(I);
        std::vector<int> ns(I);
        std::vector<JobSpec> jobs;
        auto xk = [&](int i, int k) { return k == 0 ? in[i].x1 : in[i].x2; };
        const char* resn[2] = {".conv1", ".conv2"};
        // x_k_ = ResidualBlock(x_k): relu(conv1) then conv2 + identity  (:61-62, :31-35)
        for (int i = 0; i < I; ++i) for (int k = 0; k < 2; ++k) {
            t[i][k] = alloc();
            JobSpec j; j.segs = {{0, xk(i, k)}}; j.weight = W(p + resn[k] + ".conv1"); j.relu = true; j.out_slot = t[i][k];
            jobs.push_back(j);
        }
        gemm(jobs, 128, 9); jobs.clear();
        for (int i = 0; i < I; ++i) for (int k = 0; k < 2; ++k) {
            r[i][k] = alloc();
            JobSpec j; j.segs = {{0, t[i][k]}}; j.weight = W(p + resn[k] + ".conv2"); j.res_slot = xk(i, k); j.out_slot = r[i][k];
            jobs.push_back(j);
        }
        gemm(jobs, 128, 9); jobs.clear();
        for (int i = 0; i < I; ++i) for (int k = 0; k < 2; ++k) release(t[i][k]);
        // y_k = norm_s(convf_k(cat[x_s, x_other]))  (:63-64), LayerNorm fused in the epilogue
        for (int i = 0; i < I; ++i) for (int k = 0; k < 2; ++k) {
            y[i][k] = alloc();
            JobSpec j; j.segs = {{0, in[i].xs}, {0, xk(i, 1 - k)}}; j.weight = W(p + (k == 0 ? ".convf1" : ".convf2"));
            j.ln = W(p + ".norm_s"); j.out_slot = y[i][k];
            jobs.push_back(j);
        }
        gemm(jobs, 128, 1); jobs.clear();
        // centres_k = clustering(y_k) ; v_k = v_k(x_k)  (:63-67)
        for (int i = 0; i < I; ++i) for (int k = 0; k < 2; ++k) {
            c[i][k] = alloc();
            JobSpec j; j.segs = {{0, y[i][k]}}; j.weight = W(p + ".clustering"); j.out_slot = c[i][k];
            jobs.push_back(j);
        }
        for (int i = 0; i < I; ++i) for (int k = 0; k < 2; ++k) {
            v[i][k] = alloc();
            JobSpec j; j.segs = {{0, xk(i, k)}}; j.weight = W(p + (k == 0 ? ".v1" : ".v2")); j.out_slot = v[i][k];
            jobs.push_back(j);
        }
        gemm(jobs, 128, 1); jobs.clear();
        for (int i = 0; i < I; ++i) for (int k = 0; k < 2; ++k) release(y[i][k]);
        // att_k = centres_k . v_k^T * nf^-0.5 ; softmax  (:69-73)
        std::vector<std::pair<int, int>> cv;
        for (int i = 0; i < I; ++i) for (int k = 0; k < 2; ++k) cv.push_back({c[i][k], v[i][k]});
        attention(cv);
        // out_k = softmax(att_k) . v_k, returned as (out_1 + x_2_, out_2 + x_1_)  (:72-77)
        for (int i = 0; i < I; ++i) for (int k = 0; k < 2; ++k) {
            nx[i][k] = alloc();
            JobSpec j; j.segs = {{0, v[i][k]}}; j.dyn_pair = i * 2 + k; j.res_slot = r[i][1 - k]; j.out_slot = nx[i][k];
            jobs.push_back(j);
        }
        gemm(jobs, 128, 1); jobs.clear();
        // x_s_ = unclustering(cat[centres_1, centres_2]) + x_s  (:75)
        for (int i = 0; i < I; ++i) {
            ns[i] = alloc();
            JobSpec j; j.segs = {{0, c[i][0]}, {0, c[i][1]}}; j.weight = W(p + ".unclustering"); j.res_slot = in[i].xs; j.out_slot = ns[i];
            jobs.push_back(j);
        }
        gemm(jobs, 128, 1); jobs.clear();
        std::vector<Tri> out(I);
        for (int i = 0; i < I; ++i) {
            for (int k = 0; k < 2; ++k) { release(c[i][k]); release(v[i][k]); release(r[i][k]); }
            release(in[i].x1); release(in[i].x2); release(in[i].xs);
            out[i] = {nx[i][0], nx[i][1], ns[i]};
        }
        return out;
    }

    // BIE.forward with the 1x1 / attention section fused (bie_fused.cu): conv1 of the two
    // ResidualBlocks, bie_front_tc (x_s' and the attention partial sums), att_fold (per-image mix
    // matrices), then conv2 of the ResidualBlocks with the mix segment:
    //   x_k' = out_k + x_{1-k}_ = M_k[b] x_k + P_k bv_k + x_{1-k} + conv2(relu(conv1(x_{1-k})))
    std::vector<Tri> bie_fused(const std::string& p, const std::vector<Tri>& in) {
        const int I = (int)in.size();
        std::vector<std::array<int, 2>> t(I), nx(I);
        std::vector<int> ns(I);
        std::vector<JobSpec> jobs;
        auto xk = [&](int i, int k) { return k == 0 ? in[i].x1 : in[i].x2; };
        const char* resn[2] = {".conv1", ".conv2"};
        for (int i = 0; i < I; ++i) for (int k = 0; k < 2; ++k) {
            t[i][k] = alloc();
            JobSpec j; j.segs = {{0, xk(i, k)}}; j.weight = W(p + resn[k] + ".conv1"); j.relu = true; j.out_slot = t[i][k];
            jobs.push_back(j);
        }
        gemm(jobs, 128, 9); jobs.clear();
        for (int i = 0; i < I; ++i) ns[i] = in[i].xs;           // x_s' is accumulated onto x_s in place (bie_front_tc)
        if (!m->dry) {
            const Geom& g = m->g;
            Op f;
            f.kind = Op::kBieFront;
            memset(&f.fp, 0, sizeof(f.fp));
            BieFrontParams& q = f.fp;
            q.map_act = m->map_act; q.map_w = m->map_w128; q.map_out = m->map_act_out;
            q.n_inst = I;
            for (int i = 0; i < I; ++i) {
                q.inst[i].x1_row = (int)(in[i].x1 * g.rows()); q.inst[i].x2_row = (int)(in[i].x2 * g.rows());
                q.inst[i].xs_row = (int)(in[i].xs * g.rows()); q.inst[i].out_row = (int)(ns[i] * g.rows());
            }
            const WeightSpec &wf = m->weights[W(p + ".convf1")], &wc = m->weights[W(p + ".clustering_ln")],
                             &wu = m->weights[W(p + ".unclustering")];
            q.wf_row = wf.row_base; q.wc_row = wc.row_base; q.wu_row = wu.row_base;
            q.bf = m->f32_dev + wf.bias_off; q.bc = m->f32_dev + wc.bias_off; q.bu = m->f32_dev + wu.bias_off;
            q.ln_eps = 1e-6f;                                             // norm_s (submodules.py:48: LayerNorm2d default eps)
            q.act_base = m->slot_ptr(0); q.out_base = m->slot_ptr(0);
            q.g_partial = m->gpart_ptr(); q.s_partial = m->spart_ptr();
            q.g = g; q.tiles_per_img = g.R / 128;
            q.total_tiles = I * g.B * q.tiles_per_img;
            const int grid = bie_front_grid(q.total_tiles);
            q.tiles_per_cta = (q.total_tiles + grid - 1) / grid;
            q.lcm = bie_front_lcm(q.tiles_per_cta, q.tiles_per_img);
            m->ops.push_back(f);
            Op d;
            d.kind = Op::kFold;
            memset(&d.dp, 0, sizeof(d.dp));
            FoldParams& fo = d.dp;
            fo.g_partial = q.g_partial; fo.s_partial = q.s_partial;
            fo.n_inst = I; fo.B = g.B; fo.tiles_per_img = q.tiles_per_img; fo.tiles_per_cta = q.tiles_per_cta;
            fo.total_tiles = q.total_tiles; fo.lcm = q.lcm;
            fo.w_base = m->w_dev;
            const WeightSpec &wv1 = m->weights[W(p + ".v1")], &wv2 = m->weights[W(p + ".v2")];
            fo.wv_row[0] = wv1.row_base; fo.wv_row[1] = wv2.row_base;
            fo.bv[0] = m->f32_dev + wv1.bias_off; fo.bv[1] = m->f32_dev + wv2.bias_off;
            fo.scale = 1.0f / sqrtf(128.f);                              // nf ** -0.5, submodules.py:47
            fo.m_base = m->p_ptr(); fo.bias_img = m->bimg_ptr();
            m->ops.push_back(d);
        }
        for (int i = 0; i < I; ++i) for (int k = 0; k < 2; ++k) {
            nx[i][k] = alloc();
            JobSpec j; j.segs = {{0, t[i][1 - k]}}; j.weight = W(p + resn[1 - k] + ".conv2");
            j.mix_slot = xk(i, k); j.mix_pair = i * 2 + k; j.ident_slot = xk(i, 1 - k); j.out_slot = nx[i][k];
            jobs.push_back(j);
        }
        gemm(jobs, 128, 9); jobs.clear();
        std::vector<Tri> out(I);
        for (int i = 0; i < I; ++i) {
            for (int k = 0; k < 2; ++k) release(t[i][k]);
            release(in[i].x1); release(in[i].x2);               // (the x_s slot lives on as x_s')
            out[i] = {nx[i][0], nx[i][1], ns[i]};
        }
        return out;
    }

    void build_plain() {
        const std::string blk = "neuro.para_reschunk.";
        const int H = m->slot_h[0];
        int x1 = alloc(), x2 = alloc(), xs = alloc();
        std::vector<JobSpec> jobs(3);
        // BMCNet_plain.py:27-29
        jobs[0].segs = {{0, H}, {1, 0}}; jobs[0].weight = W("f1"); jobs[0].relu = true; jobs[0].out_slot = x1;
        jobs[1].segs = {{0, H}, {1, 0}}; jobs[1].weight = W("f2"); jobs[1].relu = true; jobs[1].out_slot = x2;
        jobs[2].segs = {{0, H}, {1, 0}}; jobs[2].weight = W("fs"); jobs[2].relu = true; jobs[2].out_slot = xs;
        gemm(jobs, 128, 9);
        for (int i = 0; i < m->n_b; ++i) {
            auto o = bie(blk + std::to_string(i), {{x1, x2, xs}});
            x1 = o[0].x1; x2 = o[0].x2; xs = o[0].xs;
        }
        // x_h = relu(conv_h(xs)) written in place of the consumed hidden state; x_o = conv_o(cat[x1,x2])
        JobSpec jh; jh.segs = {{0, xs}}; jh.weight = W("h"); jh.relu = true; jh.out_slot = H; jh.hid32 = 0;
        gemm({jh}, 128, 9);
        JobSpec jo; jo.segs = {{0, x1}, {0, x2}}; jo.weight = W("o"); jo.out_f32 = true;
        gemm({jo}, 32, 9);
        release(x1); release(x2); release(xs);
    }

    void build_full() {
        const std::string blk = "neuro.para_reschunk.";
        // Backbone.forward(xs, hp, hn, hs, o) is CALLED with (x_h, x_h_p, x_h_n) (BMCNet.py:57 vs
        // :115): slot_h[0] holds x_h and plays hp, slot_h[1] = x_h_p plays hn, slot_h[2] = x_h_n plays hs.
        const int HP = m->slot_h[0], HN = m->slot_h[1], HS = m->slot_h[2];
        int xp_st = alloc(), xn_st = alloc(), xp_s = alloc(), xn_s = alloc();
        std::vector<JobSpec> jobs(4);
        jobs[0].segs = {{0, HP}, {1, 0}}; jobs[0].weight = W("fpst"); jobs[0].out_slot = xp_st;   // :64
        jobs[1].segs = {{0, HN}, {1, 0}}; jobs[1].weight = W("fnst"); jobs[1].out_slot = xn_st;   // :65
        jobs[2].segs = {{0, HP}, {1, 0}}; jobs[2].weight = W("fps"); jobs[2].out_slot = xp_s;     // :66
        jobs[3].segs = {{0, HN}, {1, 0}}; jobs[3].weight = W("fns"); jobs[3].out_slot = xn_s;     // :67
        for (auto& j : jobs) j.relu = true;
        gemm(jobs, 128, 9);
        int xs = alloc(), xs_p = alloc(), xs_n = alloc();
        jobs.assign(3, JobSpec());
        const int hsel[3] = {HS, HP, HN}, outs[3] = {xs, xs_p, xs_n};                            // :70-73
        for (int j = 0; j < 3; ++j) {
            jobs[j].segs = {{0, xp_st}, {0, xn_st}, {0, hsel[j]}, {1, 0}};
            jobs[j].weight = W("fs"); jobs[j].relu = true; jobs[j].out_slot = outs[j];
        }
        gemm(jobs, 128, 9);
        for (int i = 0; i < m->n_b; ++i) {
            const std::string p = blk + std::to_string(i);
            // ParallelBlk.forward (BMCNet.py:19-32): four ResidualBlocks ...
            const int xin[4] = {xp_s, xn_s, xp_st, xn_st};
            const char* nm[4] = {".conv1", ".conv2", ".conv1_st", ".conv2_st"};
            int t[4], xo[4];
            jobs.assign(4, JobSpec());
            for (int j = 0; j < 4; ++j) {
                t[j] = alloc();
                jobs[j].segs = {{0, xin[j]}}; jobs[j].weight = W(p + nm[j] + ".conv1"); jobs[j].relu = true; jobs[j].out_slot = t[j];
            }
            gemm(jobs, 128, 9);
            jobs.assign(4, JobSpec());
            for (int j = 0; j < 4; ++j) {
                jobs[j].segs = {{0, t[j]}}; jobs[j].weight = W(p + nm[j] + ".conv2");
                if (m->fused && m->inplace_res) {
                    // x + conv2(relu(conv1(x))) IN PLACE: the launch is a plain 3x3 convolution of t whose epilogue adds
                    // the result onto x in its own slot (TMA reduce-add): no identity K segment (it cost 59 us per
                    // launch on conv_slabt_tc), the launch runs on the two-tile kernel
                    xo[j] = xin[j];
                    jobs[j].out_slot = xo[j]; jobs[j].accumulate = true;
                } else {
                    xo[j] = alloc();
                    jobs[j].out_slot = xo[j];
                    if (m->fused) jobs[j].ident_slot = xin[j]; else jobs[j].res_slot = xin[j];
                }
            }
            gemm(jobs, 128, 9);
            for (int j = 0; j < 4; ++j) { release(t[j]); if (xo[j] != xin[j]) release(xin[j]); }
            // ... local BIE on each polarity (shared weights), then the global BIE across them
            auto l = bie(p + ".lBIE", {{xo[0], xo[2], xs_p}, {xo[1], xo[3], xs_n}});
            xp_st = l[0].x2; xs_p = l[0].xs; xn_st = l[1].x2; xs_n = l[1].xs;
            auto gl = bie(p + ".gBIE", {{l[0].x1, l[1].x1, xs}});
            xp_s = gl[0].x1; xn_s = gl[0].x2; xs = gl[0].xs;
        }
        // BMCNet.py:78-82.  Each new hidden state overwrites the slot it is read from next step.
        jobs.assign(3, JobSpec());
        const int hin[3] = {xs, xs_p, xs_n};
        const char* hn[3] = {"conv_hs", "conv_hp", "conv_hn"};
        for (int j = 0; j < 3; ++j) {
            jobs[j].segs = {{0, hin[j]}}; jobs[j].weight = W(hn[j]); jobs[j].relu = true; jobs[j].out_slot = m->slot_h[j];
            jobs[j].hid32 = j;
        }
        gemm(jobs, 128, 9);
        JobSpec jo; jo.segs = {{0, xp_s}, {0, xn_s}}; jo.weight = W("o"); jo.out_f32 = true;
        gemm({jo}, 32, 9);
        for (int s : {xp_s, xn_s, xs, xp_st, xn_st, xs_p, xs_n}) release(s);
    }

    void build() {
        m->ops.clear(); m->free_slots.clear(); m->slots_hi = 0;
        const int nh = m->kind == BMC_MODEL_BMCNET ? 3 : 1;
        for (int i = 0; i < nh; ++i) m->slot_h[i] = alloc();
        if (m->kind == BMC_MODEL_BMCNET) build_full(); else build_plain();
    }
};

void drop_graph(bmc_model* m) {
    if (m->graph_exec) { cudaGraphExecDestroy(m->graph_exec); m->graph_exec = nullptr; }
    if (m->graph) { cudaGraphDestroy(m->graph); m->graph = nullptr; }
    m->eager_runs = 0;
}

int launch_op(bmc_model* m, const Op& op, cudaStream_t st);

// BMC_OP_TIMES=1: time every launch of the plan with CUDA events (warm caches, back to back) and print the
// list -- the in-context complement of an ncu launch list, whose replays start from a flushed L2.
int time_ops(bmc_model* m, cudaStream_t st) {
    const char* names[] = {"conv_gemm", "att", "att_softmax", "bie_front", "att_fold"};
    std::vector<cudaEvent_t> ev(m->ops.size() + 1);
    for (auto& e : ev) cudaEventCreate(&e);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(ev[0], st);
        for (size_t i = 0; i < m->ops.size(); ++i) {
            int rc = launch_op(m, m->ops[i], st);
            if (rc) return rc;
            cudaEventRecord(ev[i + 1], st);
        }
    }
    BMC_CUDA(cudaStreamSynchronize(st));
    float total = 0.f;
    for (size_t i = 0; i < m->ops.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
        const Op& op = m->ops[i];
        if (op.kind == Op::kGemm)
            printf("optime %3zu %-12s jobs %d segs %d taps %d n %3d mix %d res %d : %8.1f us\n", i, names[op.kind], op.gp.n_jobs,
                   op.gp.n_seg, op.gp.n_taps, op.gp.n, op.gp.tap1_mask, op.gp.jobs[0].residual != nullptr, ms * 1e3f);
        else
            printf("optime %3zu %-12s : %8.1f us\n", i, names[op.kind], ms * 1e3f);
        total += ms;
    }
    printf("optime total %.1f us over %zu launches\n", total * 1e3f, m->ops.size());
    for (auto& e : ev) cudaEventDestroy(e);
    return BMC_OK;
}

int launch_op(bmc_model* m, const Op& op, cudaStream_t st) {
    if (op.kind == Op::kGemm) return launch_conv_gemm(op.gp, m->simt, st);
    if (op.kind == Op::kAtt) return launch_att(op.ap, m->simt, st);
    if (op.kind == Op::kBieFront) return launch_bie_front(op.fp, st);
    if (op.kind == Op::kFold) {
        static int simt_fold = -1;
        if (simt_fold < 0) simt_fold = measure_env("BMC_FOLD_SIMT", 0);
        return simt_fold ? launch_att_fold(op.dp, st) : launch_att_fold_tc(op.dp, m->map_w128, st);
    }
    return launch_att_softmax(op.sp, st);
}

int run_ops(bmc_model* m, cudaStream_t st) {
    static int want_times = -1;
    if (want_times < 0) want_times = measure_env("BMC_OP_TIMES", 0);
    static int calls = 0;
    if (want_times && ++calls == 4) { want_times = 0; return time_ops(m, st); }
    for (const Op& op : m->ops) {
        int rc = BMC_OK;
        if (op.kind == Op::kGemm) rc = launch_conv_gemm(op.gp, m->simt, st);
        else if (op.kind == Op::kAtt) rc = launch_att(op.ap, m->simt, st);
        else if (op.kind == Op::kBieFront) rc = launch_bie_front(op.fp, st);
        else if (op.kind == Op::kFold) rc = launch_op(m, op, st);
        else rc = launch_att_softmax(op.sp, st);
        if (rc) return rc;
    }
    return BMC_OK;
}

// Replay the step as one CUDA graph (the ops only touch the arena and the weights, so the
// captured kernel arguments never change).  The first call runs eagerly (it also sets the
// kernels' shared-memory attributes, which must not happen inside a capture).
int run_plan(bmc_model* m, cudaStream_t st) {
    if (!m->use_graph) return run_ops(m, st);
    if (m->graph_exec && m->graph_simt != m->simt) drop_graph(m);
    if (!m->graph_exec) {
        if (m->eager_runs++ == 0) return run_ops(m, st);
        if (!m->cap_stream) BMC_CUDA(cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking));
        BMC_CUDA(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
        int rc = run_ops(m, m->cap_stream);
        cudaError_t e = cudaStreamEndCapture(m->cap_stream, &m->graph);
        if (rc) return rc;
        BMC_CUDA(e);
        BMC_CUDA(cudaGraphInstantiate(&m->graph_exec, m->graph, 0));
        m->graph_simt = m->simt;
    }
    BMC_CUDA(cudaGraphLaunch(m->graph_exec, st));
    return BMC_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------- C ABI
extern "C" BMC_EXPORT bmc_model_t* bmc_model_create(int kind, int scale, int n_c, int n_b, int repeat) {
    if ((kind != BMC_MODEL_BMCNET && kind != BMC_MODEL_BMCNET_PLAIN) || scale != 4 || n_c != 128 || repeat != 3 ||
        n_b < 1 || n_b > 64) {
        set_error("bmc_model_create: kernels are specialised for scale=4, n_c=128, repeat=3 (got kind=%d scale=%d "
                  "n_c=%d n_b=%d repeat=%d)", kind, scale, n_c, n_b, repeat);
        return nullptr;
    }
    bmc_model* m = new bmc_model();
    m->kind = kind; m->scale = scale; m->n_c = n_c; m->n_b = n_b; m->repeat = repeat;
    m->use_graph = !measure_env("BMC_NO_GRAPH", 0);
    register_weights(m);
    return m;
}

extern "C" BMC_EXPORT void bmc_model_destroy(bmc_model_t* m) {
    if (!m) return;
    drop_graph(m);
    if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
    delete m;
}

extern "C" BMC_EXPORT size_t bmc_model_weight_bytes(const bmc_model_t* m) {
    return align_up((size_t)m->w_rows_total * 128, 1024) + align_up(m->f32_floats * 4, 1024) +
           align_up(m->kmap_ints * 4, 1024);
}

extern "C" BMC_EXPORT int bmc_model_load_state_dict(bmc_model_t* m, const char* const* names, const float* const* tensors,
                                         const int64_t* numels, int n_tensors, void* weight_buf,
                                         size_t weight_buf_bytes, void* stream) {
    BMC_REQUIRE(m && names && tensors && numels && weight_buf, "load_state_dict: NULL argument");
    BMC_REQUIRE(weight_buf_bytes >= bmc_model_weight_bytes(m), "load_state_dict: weight buffer too small (%zu < %zu)",
                weight_buf_bytes, bmc_model_weight_bytes(m));
    BMC_REQUIRE(((uintptr_t)weight_buf & 1023) == 0, "load_state_dict: weight buffer must be 1024-byte aligned");
    cudaStream_t st = as_stream(stream);
    // strict=True semantics (infer_BMCNet.py:112): exactly the reference key set
    const auto keys = canonical_keys(m->kind, m->n_b);
    std::map<std::string, int> given;
    for (int i = 0; i < n_tensors; ++i) given[names[i]] = i;
    for (auto& k : keys)
        BMC_REQUIRE(given.count(k), "load_state_dict: missing key \"%s\"", k.c_str());
    for (auto& kv : given)
        BMC_REQUIRE(std::find(keys.begin(), keys.end(), kv.first) != keys.end(), "load_state_dict: unexpected key \"%s\"",
                    kv.first.c_str());
    // aliased keys write the same Parameter in the reference; the last one in module order wins
    std::map<std::string, std::string> last;
    for (auto& k : keys) last[alias_root(k)] = k;
    auto find = [&](const std::string& key, int64_t want, const float** out) -> int {
        const std::string eff = last[alias_root(key)];
        const int i = given[eff];
        if (numels[i] != want) {
            set_error("load_state_dict: size mismatch for \"%s\": %lld elements, expected %lld", eff.c_str(),
                      (long long)numels[i], (long long)want);
            return BMC_ERR_ARG;
        }
        *out = tensors[i];
        return BMC_OK;
    };
    char* base = static_cast<char*>(weight_buf);
    m->w_dev = reinterpret_cast<act_t*>(base);
    m->f32_dev = reinterpret_cast<float*>(base + align_up((size_t)m->w_rows_total * 128, 1024));
    m->kmap_dev = reinterpret_cast<int*>(reinterpret_cast<char*>(m->f32_dev) + align_up(m->f32_floats * 4, 1024));
    BMC_CUDA(cudaMemsetAsync(m->f32_dev, 0, m->f32_floats * 4, st));
    std::vector<int> kmap_host(m->kmap_ints);
    size_t koff = 0;
    std::vector<size_t> koffs;
    for (auto& w : m->weights) {
        koffs.push_back(koff);
        for (auto& s : w.segs)
            for (int t = 0; t < w.taps; ++t)
                for (int c : s.src) kmap_host[koff++] = c < 0 ? -1 : c * w.taps + t;
    }
    BMC_CUDA(cudaMemcpyAsync(m->kmap_dev, kmap_host.data(), m->kmap_ints * 4, cudaMemcpyHostToDevice, st));
    BMC_CUDA(cudaStreamSynchronize(st));      // kmap_host is a temporary
    for (size_t i = 0; i < m->weights.size(); ++i) {
        const WeightSpec& w = m->weights[i];
        const float *wt = nullptr, *bs = nullptr;
        int rc = find(w.conv + ".weight", (int64_t)w.n_out * w.cin * w.taps, &wt);
        if (rc) return rc;
        rc = find(w.conv + ".bias", w.n_out, &bs);
        if (rc) return rc;
        // weight w occupies rows [row_base, row_base + k_chunks * n_out): chunk-major [k_chunks][n_out][64]
        const float *gam = nullptr, *bet = nullptr;
        if (!w.fold_ln.empty()) {
            rc = find(w.fold_ln + ".weight", w.cin, &gam);
            if (rc) return rc;
            rc = find(w.fold_ln + ".bias", w.cin, &bet);
            if (rc) return rc;
        }
        rc = launch_repack_weight(wt, m->kmap_dev + koffs[i], w.cin * w.taps, w.n_out, w.n_out, w.k_chunks * 64,
                                  m->w_dev + (size_t)w.row_base * 64, w.n_out, 0, st, gam);
        if (rc) return rc;
        if (bet) {
            rc = launch_fold_beta_bias(wt, bs, bet, w.n_out, w.cin, m->f32_dev + w.bias_off, st);
            if (rc) return rc;
        } else {
            BMC_CUDA(cudaMemcpyAsync(m->f32_dev + w.bias_off, bs, w.n_out * 4, cudaMemcpyDeviceToDevice, st));
        }
    }
    {
        int rc = launch_fill_identity(m->w_dev + (size_t)m->ident_row * 64, st);
        if (rc) return rc;
    }
    for (auto& ln : m->lns) {
        const float *g = nullptr, *b = nullptr;
        int rc = find(ln.prefix + ".weight", 128, &g);
        if (rc) return rc;
        rc = find(ln.prefix + ".bias", 128, &b);
        if (rc) return rc;
        BMC_CUDA(cudaMemcpyAsync(m->f32_dev + ln.gamma_off, g, 512, cudaMemcpyDeviceToDevice, st));
        BMC_CUDA(cudaMemcpyAsync(m->f32_dev + ln.beta_off, b, 512, cudaMemcpyDeviceToDevice, st));
    }
    int rc = make_tmap_2d_act(&m->map_w128, m->w_dev, (uint64_t)m->w_rows_total, 64, 128, 64);
    if (rc) return rc;
    rc = make_tmap_2d_act(&m->map_w64, m->w_dev, (uint64_t)m->w_rows_total, 64, 64, 64);
    if (rc) return rc;
    rc = make_tmap_2d_act(&m->map_w32, m->w_dev, (uint64_t)m->w_rows_total, 64, 32, 64);
    if (rc) return rc;
    rc = make_tmap_2d_act_sw64(&m->map_w_s64, m->w_dev, (uint64_t)m->w_rows_total, 64, 128);
    if (rc) return rc;
    m->loaded = true;
    if (m->bound) {            // weights moved: rebuild the plan against the new pointers
        drop_graph(m);
        m->dry = false;
        Builder b{m};
        b.build();
    }
    return BMC_OK;
}

extern "C" BMC_EXPORT int bmc_model_configure(bmc_model_t* m, int batch, int H, int W) {
    BMC_REQUIRE(m, "configure: NULL model");
    BMC_REQUIRE(batch >= 1 && H >= 1 && W >= 1, "configure: bad shape B=%d H=%d W=%d", batch, H, W);
    m->g = Geom::make(batch, H, W);
    BMC_REQUIRE(m->g.rows() * 64 < (1L << 31), "configure: batch x image too large for 32-bit row coordinates");
    drop_graph(m);
    m->bound = false;
    m->dry = true;
    // the fused BIE plan needs the slab kernel (mix segment); very wide images fall back to the unfused plan
    {
        GemmParams probe;
        memset(&probe, 0, sizeof(probe));
        probe.n = 128; probe.n_taps = 9; probe.g = m->g;
        m->allow_fused = slab_supported(probe) && measure_env("BMC_FUSED", 1);
        m->inplace_res = slab2_geom_supported(m->g);
    }
    Builder b{m};
    m->fused = false;
    b.build();
    m->n_slots = m->slots_hi;
    if (m->allow_fused) {                 // arena must hold either plan (the SIMT cross-check runs the unfused one)
        m->fused = true;
        b.build();
        m->n_slots = std::max(m->n_slots, m->slots_hi);
    }
    m->fused = m->allow_fused && !m->simt;
    const int chunks = m->g.R / 64;
    int want = (sm_count() + 4 * batch - 1) / (4 * batch);
    want = std::max(1, std::min(want, chunks));
    m->pix_per_split = (chunks + want - 1) / want * 64;
    m->n_split = (m->g.R + m->pix_per_split - 1) / m->pix_per_split;
    const size_t rows = (size_t)m->g.rows();
    size_t off = align_up((size_t)m->n_slots * rows * 256, 1024);
    m->off_mi = off; off += align_up(rows * 128, 1024);
    m->off_a32 = off; off += align_up(rows * 128, 1024);
    m->off_p = off; off += align_up((size_t)kMaxPairs * batch * 256 * 128, 1024);
    m->off_partial = off; off += align_up((size_t)kMaxPairs * batch * m->n_split * 128 * 128 * 4, 1024);
    m->off_hid32 = off; off += align_up((size_t)(m->kind == BMC_MODEL_BMCNET ? 3 : 1) * rows * 512, 1024);
    m->off_bimg = off; off += align_up((size_t)kMaxPairs * batch * 128 * 4, 1024);
    {
        const int n_inst = m->kind == BMC_MODEL_BMCNET ? 2 : 1;
        const int slots = bie_front_slots(n_inst * batch * (m->g.R / 128), m->g.R / 128);
        m->off_gpart = off; off += align_up((size_t)slots * 2 * 128 * 128 * 4, 1024);
        m->off_spart = off; off += align_up((size_t)slots * 2 * 128 * 4, 1024);
    }
    m->ws_bytes = off;
    m->configured = 1;
    return BMC_OK;
}

extern "C" BMC_EXPORT size_t bmc_model_workspace_bytes(const bmc_model_t* m) { return m && m->configured ? m->ws_bytes : 0; }

extern "C" BMC_EXPORT int bmc_model_bind_workspace(bmc_model_t* m, void* workspace, size_t workspace_bytes) {
    BMC_REQUIRE(m && m->configured, "bind_workspace: configure the model first");
    if (!m->loaded) { set_error("bind_workspace: load the weights first"); return BMC_ERR_STATE; }
    if (!workspace || workspace_bytes < m->ws_bytes) {
        set_error("bind_workspace: need %zu bytes, got %zu", m->ws_bytes, workspace_bytes);
        return BMC_ERR_WORKSPACE;
    }
    BMC_REQUIRE(((uintptr_t)workspace & 1023) == 0, "bind_workspace: workspace must be 1024-byte aligned");
    drop_graph(m);
    m->ws = static_cast<char*>(workspace);
    // halo / tail rows must be zero and are never written with anything else afterwards
    // (bind has no stream argument: clear on the legacy stream and drain the device, so that kernels later
    // launched on ANY stream -- including cudaStreamNonBlocking ones, which do not order against the legacy
    // stream -- see the zeros)
    BMC_CUDA(cudaMemset(m->ws, 0, m->ws_bytes));
    BMC_CUDA(cudaDeviceSynchronize());
    const uint64_t rows = (uint64_t)m->g.rows();
    int rc = make_tmap_2d_act(&m->map_act, m->slot_ptr(0), rows * m->n_slots, 128, 128, 64);
    if (!rc) rc = make_tmap_2d_act(&m->map_att, m->slot_ptr(0), rows * m->n_slots, 128, 64, 64);
    m->abox = slab_box_rows(m->g, 9);
    if (!rc) rc = make_tmap_2d_act(&m->map_slab, m->slot_ptr(0), rows * m->n_slots, 128, (uint32_t)m->abox, 64);
    if (!rc) rc = make_tmap_2d_act(&m->map_mi, m->mi_ptr(), rows, 64, 128, 64);
    if (!rc) rc = make_tmap_2d_act(&m->map_mi64, m->mi_ptr(), rows, 64, (uint32_t)m->abox, 64);
    if (!rc) rc = make_tmap_2d_act(&m->map_p, m->p_ptr(), (uint64_t)kMaxPairs * m->g.B * 256, 64, 128, 64);
    m->abox32 = slab2_box_rows(m->g, 9);
    if (!rc) rc = make_tmap_2d_act_sw64(&m->map_act_s64, m->slot_ptr(0), rows * m->n_slots, 128, (uint32_t)m->abox32);
    if (!rc) rc = make_tmap_2d_act_sw64(&m->map_mi_s64, m->mi_ptr(), rows, 64, (uint32_t)m->abox32);
    if (!rc) rc = make_tmap_2d_act_sw64(&m->map_p_s64, m->p_ptr(), (uint64_t)kMaxPairs * m->g.B * 256, 64, 128);
    if (!rc) rc = make_tmap_2d_act(&m->map_act_out, m->slot_ptr(0), rows * m->n_slots, 128, 32, 64);
    if (rc) return rc;
    m->dry = false;
    Builder b{m};
    b.build();
    m->bound = true;
    return BMC_OK;
}

extern "C" BMC_EXPORT int bmc_model_launches_per_step(const bmc_model_t* m) { return m ? (int)m->ops.size() : 0; }

extern "C" BMC_EXPORT int bmc_model_set_debug_simt(bmc_model_t* m, int enable) {
    BMC_REQUIRE(m, "set_debug_simt: NULL model");
    m->simt = enable ? 1 : 0;
    const bool fused = m->allow_fused && !m->simt;
    if (fused != m->fused) {              // the SIMT cross-check runs the unfused dataflow: rebuild the plan
        m->fused = fused;
        if (m->bound) {
            drop_graph(m);
            m->dry = false;
            Builder b{m};
            b.build();
        }
    }
    return BMC_OK;
}

namespace {
int check_ready(bmc_model* m) {
    if (!m || !m->bound || !m->loaded) { set_error("model not ready: create -> load_state_dict -> configure -> bind_workspace"); return BMC_ERR_STATE; }
    return BMC_OK;
}
}  // namespace

extern "C" BMC_EXPORT int bmc_model_forward(bmc_model_t* m, const float* x, const int64_t x_strides[5], const float* x_h,
                                 const float* x_h_p, const float* x_h_n, const float* x_o, int init, float* out_h,
                                 float* out_h_p, float* out_h_n, float* out_o, void* stream) {
    int rc = check_ready(m);
    if (rc) return rc;
    const bool full = m->kind == BMC_MODEL_BMCNET;
    BMC_REQUIRE(x && out_h && out_o, "forward: NULL tensor");
    BMC_REQUIRE(!full || (out_h_p && out_h_n), "forward: BMCNet needs out_h_p / out_h_n");
    BMC_REQUIRE(!full || ((x_h != nullptr) == (x_h_p != nullptr) && (x_h != nullptr) == (x_h_n != nullptr)),
                "forward: pass all hidden states or none");
    BMC_REQUIRE(x_o || !init, "forward: x_o may only be omitted when init == 0 (fed-back prediction kept on the device)");
    cudaStream_t st = as_stream(stream);
    const float* hin[3] = {x_h, x_h_p, x_h_n};
    float* hout[3] = {out_h, out_h_p, out_h_n};
    const int nh = full ? 3 : 1;
    for (int i = 0; i < nh && x_h; ++i) {          // x_h == NULL: the states of the previous call are still resident
        rc = launch_pack_nchw(hin[i], m->g, 128, m->slot_ptr(m->slot_h[i]), 128, 0, st);
        if (rc) return rc;
    }
    PackInputsParams pi;
    pi.x = x; for (int i = 0; i < 5; ++i) pi.xs[i] = x_strides[i];
    pi.x_o = x_o; pi.init = init; pi.mi = m->mi_ptr(); pi.g = m->g;
    rc = launch_pack_inputs(pi, st);
    if (rc) return rc;
    rc = run_plan(m, st);
    if (rc) return rc;
    for (int i = 0; i < nh; ++i) {
        rc = launch_unpack_nchw(m->slot_ptr(m->slot_h[i]), m->g, 128, 128, 0, hout[i], st);   // the 16-bit state, widened
        if (rc) return rc;
    }
    EmitParams ep;
    ep.a = m->a32_ptr(); ep.x = x; for (int i = 0; i < 5; ++i) ep.xs[i] = x_strides[i];
    ep.out_o = out_o; ep.mi_next = m->mi_ptr(); ep.g = m->g;
    return launch_emit(ep, st);
}

extern "C" BMC_EXPORT int bmc_model_step(bmc_model_t* m, const float* x, const int64_t x_strides[5], int reset, float* out_o,
                              void* stream) {
    int rc = check_ready(m);
    if (rc) return rc;
    BMC_REQUIRE(x, "step: NULL input");
    cudaStream_t st = as_stream(stream);
    if (reset) {
        const int nh = m->kind == BMC_MODEL_BMCNET ? 3 : 1;
        for (int i = 0; i < nh; ++i)
            BMC_CUDA(cudaMemsetAsync(m->slot_ptr(m->slot_h[i]), 0, (size_t)m->g.rows() * 256, st));
    }
    PackInputsParams pi;
    pi.x = x; for (int i = 0; i < 5; ++i) pi.xs[i] = x_strides[i];
    pi.x_o = nullptr; pi.init = reset; pi.mi = m->mi_ptr(); pi.g = m->g;
    rc = launch_pack_inputs(pi, st);
    if (rc) return rc;
    rc = run_plan(m, st);
    if (rc) return rc;
    EmitParams ep;
    ep.a = m->a32_ptr(); ep.x = x; for (int i = 0; i < 5; ++i) ep.xs[i] = x_strides[i];
    ep.out_o = out_o; ep.mi_next = m->mi_ptr(); ep.g = m->g;
    return launch_emit(ep, st);
}

// ---------------------------------------------------------------------------------- per-kernel entries
extern "C" BMC_EXPORT int bmc_conv_gemm(const bmc_gemm_job_t* jobs, int n_jobs, int n, int taps, int B, int H, int W, int impl,
                             void* stream) {
    BMC_REQUIRE(jobs && n_jobs >= 1 && n_jobs <= kMaxJobs, "conv_gemm: 1..%d jobs per call", kMaxJobs);
    BMC_REQUIRE(n == 128 || n == 32, "conv_gemm: n must be 128 or 32");
    BMC_REQUIRE(taps == 1 || taps == 9, "conv_gemm: taps must be 1 or 9");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    const Geom g = Geom::make(B, H, W);
    p.n_jobs = n_jobs; p.n_seg = jobs[0].n_seg; p.n_taps = taps; p.n = n; p.g = g; p.tiles_per_img = g.R / kTileM;
    BMC_REQUIRE(p.n_seg >= 1 && p.n_seg <= 3, "conv_gemm: 1..3 segments");
    for (int t = 0; t < taps; ++t) p.tap_off[t] = taps == 9 ? (t / 3 - 1) * g.Wp + (t % 3 - 1) : 0;
    for (int s = 0; s < p.n_seg; ++s) p.chunks[s] = jobs[0].a_ch[s] / 64;
    int n_maps = 0;
    std::vector<std::pair<const void*, uint32_t>> map_key;
    auto get_map = [&](const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, int* idx) -> int {
        for (size_t i = 0; i < map_key.size(); ++i)
            if (map_key[i].first == ptr && map_key[i].second == box_rows) { *idx = (int)i; return BMC_OK; }
        if (n_maps >= kMaxMaps) { *idx = -1; return BMC_OK; }
        int rc = impl != 1 ? make_tmap_2d_act(&p.maps[n_maps], ptr, rows, cols, box_rows, 64) : BMC_OK;
        if (rc) return rc;
        map_key.push_back({ptr, box_rows});
        *idx = n_maps++;
        return BMC_OK;
    };
    bool maps64_ok = true;
    p.abox_rows = slab_box_rows(g, taps);
    for (int j = 0; j < n_jobs; ++j) {
        const bmc_gemm_job_t& js = jobs[j];
        GemmJobDev& d = p.jobs[j];
        BMC_REQUIRE(js.n_seg == p.n_seg, "conv_gemm: all jobs of a call must share the shape");
        int kch = 0;
        for (int s = 0; s < p.n_seg; ++s) {
            BMC_REQUIRE(js.a[s] && js.a_ch[s] % 64 == 0 && js.a_ch[s] / 64 == p.chunks[s], "conv_gemm: bad segment %d", s);
            int rc = get_map(js.a[s], (uint64_t)js.a_rows[s], (uint64_t)js.a_ch[s], 128, &d.a_map[s]);
            if (rc) return rc;
            BMC_REQUIRE(d.a_map[s] >= 0, "conv_gemm: more than %d distinct operand tensors in one call", kMaxMaps);
            d.a_row_base[s] = js.a_row_base[s]; d.a_col_base[s] = 0;
            d.a_ptr[s] = static_cast<const act_t*>(js.a[s]); d.a_ld[s] = js.a_ch[s]; d.a_rows[s] = js.a_rows[s];
            kch += p.chunks[s] * taps;
        }
        BMC_REQUIRE(js.w && js.w_k == kch * 64, "conv_gemm: weight K (%d) != %d", js.w_k, kch * 64);
        const uint64_t w_total = (uint64_t)js.w_row_base + (uint64_t)(B - 1) * js.w_img_stride + (uint64_t)js.w_rows * kch;
        int rc = get_map(js.w, w_total, 64, (uint32_t)n, &d.w_map);
        if (rc) return rc;
        BMC_REQUIRE(d.w_map >= 0, "conv_gemm: more than %d distinct operand tensors in one call", kMaxMaps);
        d.w_map64 = 0;
        if (n == 128 && js.w_img_stride == 0) {                     // pair kernel: half-tile boxes of the same weights
            int idx = -1;
            rc = get_map(js.w, w_total, 64, 64, &idx);
            if (rc) return rc;
            d.w_map64 = idx > 0 ? idx : 0;
        }
        d.w_ptr = static_cast<const act_t*>(js.w); d.w_rows = js.w_rows;
        d.w_row_base = js.w_row_base; d.w_img_stride = js.w_img_stride;
        d.bias = js.bias; d.relu = js.relu;
        d.residual = static_cast<const act_t*>(js.residual); d.res_row_base = js.res_row_base;
        d.out = static_cast<act_t*>(js.out_act16); d.out_row_base = js.out_row_base; d.out_f32 = js.out_f32;
        d.ln_gamma = js.ln_gamma; d.ln_beta = js.ln_beta; d.ln_eps = js.ln_eps;
    }
    // 64-row-box descriptors for the slab kernel, if they still fit; otherwise the per-tap kernel runs
    for (int j = 0; j < n_jobs && maps64_ok; ++j)
        for (int s = 0; s < p.n_seg && maps64_ok; ++s) {
            int rc = get_map(jobs[j].a[s], (uint64_t)jobs[j].a_rows[s], (uint64_t)jobs[j].a_ch[s], (uint32_t)p.abox_rows, &p.jobs[j].a_map64[s]);
            if (rc) return rc;
            if (p.jobs[j].a_map64[s] < 0) maps64_ok = false;
        }
    if (!maps64_ok)
        for (int j = 0; j < n_jobs; ++j)
            for (int s = 0; s < kMaxSeg; ++s) p.jobs[j].a_map64[s] = -1;
    // 32-channel / 64-byte-swizzle boxes for conv_slab2_tc, if the distinct tensors fit
    if (impl == 0 && n == 128) {
        int n32 = 0;
        std::vector<std::pair<const void*, uint32_t>> key32;
        bool ok = true;
        auto get32 = [&](const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, signed char* idx) -> int {
            for (size_t i = 0; i < key32.size(); ++i)
                if (key32[i].first == ptr && key32[i].second == box_rows) { *idx = (signed char)i; return BMC_OK; }
            if (n32 >= kMaxMaps32) { *idx = -1; ok = false; return BMC_OK; }
            int rc = make_tmap_2d_act_sw64(&p.maps32[n32], ptr, rows, cols, box_rows);
            if (rc) return rc;
            key32.push_back({ptr, box_rows});
            *idx = (signed char)n32++;
            return BMC_OK;
        };
        p.abox32_rows = slab2_box_rows(g, taps);
        for (int j = 0; j < n_jobs && ok; ++j) {
            GemmJobDev& d = p.jobs[j];
            for (int s = 0; s < p.n_seg && ok; ++s) {
                int rc = get32(jobs[j].a[s], (uint64_t)jobs[j].a_rows[s], (uint64_t)jobs[j].a_ch[s], (uint32_t)p.abox32_rows, &d.a_map32[s]);
                if (rc) return rc;
            }
            if (ok && jobs[j].w_img_stride == 0) {
                int kch = 0;
                for (int s = 0; s < p.n_seg; ++s) kch += p.chunks[s] * taps;
                int rc = get32(jobs[j].w, (uint64_t)jobs[j].w_row_base + (uint64_t)jobs[j].w_rows * kch, 64, 128, &d.w_map32);
                if (rc) return rc;
            } else {
                d.w_map32 = -1;
            }
        }
        p.has32 = ok ? 1 : 0;
        // output tensors behind [32 x 64] boxes for the TMA-store epilogue (rows: what this call may write)
        for (int j = 0; j < n_jobs; ++j) { p.jobs[j].out_map32 = -1; p.jobs[j].out_map_row = 0; }
        for (int j = 0; j < n_jobs && ok; ++j) {
            if (!jobs[j].out_act16) continue;
            uint64_t rows = 0;
            for (int i = 0; i < n_jobs; ++i)
                if (jobs[i].out_act16 == jobs[j].out_act16) rows = std::max<uint64_t>(rows, (uint64_t)jobs[i].out_row_base + (uint64_t)g.rows());
            int found = -1;
            for (int i = 0; i < j; ++i)
                if (jobs[i].out_act16 == jobs[j].out_act16) found = p.jobs[i].out_map32;
            if (found < 0 && n32 < kMaxMaps32) {
                int rc = make_tmap_2d_act(&p.maps32[n32], jobs[j].out_act16, rows, 128, 32, 64);
                if (rc) return rc;
                found = n32++;
            }
            p.jobs[j].out_map32 = (signed char)found;
        }
    }
    return launch_conv_gemm(p, impl, as_stream(stream));
}

extern "C" BMC_EXPORT int bmc_attention_weights(const void* centres, const void* v, int B, int H, int W, float scale,
                                     float* partial, int n_split, void* probs_act16, int impl, void* stream) {
    BMC_REQUIRE(centres && v && partial && probs_act16 && n_split >= 1, "attention_weights: bad argument");
    const Geom g = Geom::make(B, H, W);
    AttParams a;
    memset(&a, 0, sizeof(a));
    if (impl == 0) {
        int rc = make_tmap_2d_act(&a.map_c, centres, (uint64_t)g.rows(), 128, 64, 64);
        if (!rc) rc = make_tmap_2d_act(&a.map_v, v, (uint64_t)g.rows(), 128, 64, 64);
        if (rc) return rc;
    }
    a.c_ptr = static_cast<const act_t*>(centres); a.v_ptr = static_cast<const act_t*>(v);
    a.n_pairs = 1; a.n_split = n_split;
    a.pix_per_split = ((g.R / 64 + n_split - 1) / n_split) * 64;
    a.scale = scale; a.partial = partial; a.g = g;
    int rc = launch_att(a, impl, as_stream(stream));
    if (rc) return rc;
    SoftmaxParams s;
    memset(&s, 0, sizeof(s));
    s.partial = partial; s.n_pairs = 1; s.n_split = n_split; s.B = B;
    s.w_base = static_cast<act_t*>(probs_act16); s.w_img_stride = 256;
    return launch_att_softmax(s, as_stream(stream));
}

extern "C" BMC_EXPORT int bmc_layernorm_rows(const void* in_act16, const float* gamma, const float* beta, float eps, int64_t rows,
                                  void* out_act16, void* stream) {
    BMC_REQUIRE(in_act16 && gamma && beta && out_act16 && rows >= 0, "layernorm_rows: bad argument");
    return launch_layernorm(static_cast<const act_t*>(in_act16), gamma, beta, eps, rows,
                            static_cast<act_t*>(out_act16), as_stream(stream));
}

extern "C" BMC_EXPORT int bmc_pack_nchw(const float* src, int B, int C, int H, int W, void* dst_act16, int c_pad, int c_off,
                             void* stream) {
    BMC_REQUIRE(src && dst_act16 && C >= 1 && c_off >= 0 && c_off + C <= c_pad, "pack_nchw: bad argument");
    return launch_pack_nchw(src, Geom::make(B, H, W), C, static_cast<act_t*>(dst_act16), c_pad, c_off, as_stream(stream));
}

extern "C" BMC_EXPORT int bmc_unpack_nchw(const void* src_act16, int B, int C, int H, int W, int c_pad, int c_off, float* dst,
                               void* stream) {
    BMC_REQUIRE(src_act16 && dst && C >= 1 && c_off >= 0 && c_off + C <= c_pad, "unpack_nchw: bad argument");
    return launch_unpack_nchw(static_cast<const act_t*>(src_act16), Geom::make(B, H, W), C, c_pad, c_off, dst, as_stream(stream));
}
