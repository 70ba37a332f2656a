// host_pipeline.inl -- b200yolo_decode_nms_host_ws / b200yolo_decode_nms_host: the reference-facing call with
// HOST buffers (what inference.py:121 / train.py:366 hand to the model is host
// data; bench.py times this as "e2e").  Image chunks are pushed through
// kSlots streams: H2D(head0, head1) -> fused kernel -> D2H(counts) -> D2H(kept rows), so the
// PCIe copies of neighbouring chunks overlap each other and the kernel.  Only the rows an image can have kept
// travel back: one strided copy per chunk whose width is the chunk's largest count.
// Device staging is the CALLER's (b200yolo_decode_nms_host_ws; SURVEY 8b: the library never allocates);
// b200yolo_decode_nms_host is the convenience form that keeps one grow-only workspace per (thread, device).
// No lock anywhere: streams and the convenience workspace are thread-local.
// Included at the end of b200yolo.cu (uses its helpers).

namespace {

constexpr int kSlots = 3;

struct HostPlan {
    int chunk;
    bool large;
    size_t per0, per1, K;           // floats per image of head 0 / head 1; cells per image
    size_t b_h0, b_h1, b_out, b_cnt, b_ws, slot_bytes;
};

size_t up256(size_t v) { return (v + 255) / 256 * 256; }

HostPlan host_plan(int N, int A, int C, int H0, int W0, int H1, int W1) {
    HostPlan p;
    const size_t attrs = 5 + (size_t)C;
    p.per0 = (size_t)A * attrs * H0 * W0;
    p.per1 = (size_t)A * attrs * H1 * W1;
    p.K = (size_t)A * H0 * W0 + (size_t)A * H1 * W1;
    // ~4 chunks cover the batch (measured best for the 256-image batch, profiles/e2e_sweep.py), at least 8 images each
    int chunk = (N + 3) / 4;
    if (chunk < 8) chunk = 8;
    static const int env_chunk = [] { const char *e = getenv("B200YOLO_HOST_CHUNK"); return e ? atoi(e) : 0; }();
    if (env_chunk > 0) chunk = env_chunk;  // tuning knob: images per pipeline chunk
    if (chunk > N) chunk = N > 0 ? N : 1;
    p.chunk = chunk;
    // images too large for the fused kernel's shared memory go through b200yolo_decode_nms_large (records in a workspace)
    p.large = !fits_one_cta((int)p.K, C, MODE_FUSED);
    p.b_h0 = up256(p.per0 * chunk * 4);
    p.b_h1 = up256(p.per1 * chunk * 4);
    p.b_out = up256(p.K * 7 * chunk * 4);
    p.b_cnt = up256((size_t)chunk * 4);
    p.b_ws = p.large ? up256(b200yolo_decode_nms_large_workspace_bytes(chunk, (int)p.K)) : 0;
    p.slot_bytes = p.b_h0 + p.b_h1 + p.b_out + p.b_cnt + p.b_ws;
    return p;
}

struct HostStreams {
    cudaStream_t s[kSlots] = {nullptr, nullptr, nullptr};
};
thread_local HostStreams t_streams[16];
thread_local size_t t_last_d2h = 0;
struct OwnedWs { void *ptr = nullptr; size_t bytes = 0; };
thread_local OwnedWs t_owned[16];

}  // namespace

extern "C" size_t b200yolo_decode_nms_host_workspace_bytes(int N, int A, int C, int H0, int W0, int H1, int W1) {
    if (N < 1 || A < 1 || C < 1 || H0 < 1 || W0 < 1 || H1 < 1 || W1 < 1) return 0;
    return (size_t)kSlots * host_plan(N, A, C, H0, W0, H1, W1).slot_bytes + 256;
}

extern "C" size_t b200yolo_host_last_d2h_bytes(void) { return t_last_d2h; }

extern "C" int b200yolo_decode_nms_host_ws(const float *head0, const float *head1, int N, int A, int C, int H0, int W0,
                                           int H1, int W1, const float *anchor_wh, float conf_thr, double iou_thr,
                                           float *out, int *out_count, void *dev_workspace, size_t dev_workspace_bytes,
                                           int device) {
    if (!head0 || !head1 || !anchor_wh || !out || !out_count) return fail(B200YOLO_EINVAL, "decode_nms_host: null pointer");
    if (N < 0 || A < 1 || A > kMaxAnchors || C < 1 || H0 < 1 || W0 < 1 || H1 < 1 || W1 < 1 || device < 0 || device >= 16)
        return fail(B200YOLO_EINVAL, "decode_nms_host: bad argument");
    if (N == 0) return 0;
    const HostPlan hp = host_plan(N, A, C, H0, W0, H1, W1);
    const uintptr_t base = ((uintptr_t)dev_workspace + 255) & ~(uintptr_t)255;
    if (!dev_workspace || base + (size_t)kSlots * hp.slot_bytes > (uintptr_t)dev_workspace + dev_workspace_bytes)
        return fail(B200YOLO_EINVAL, "decode_nms_host: device workspace missing or too small (%zu < %zu)", dev_workspace_bytes,
                    b200yolo_decode_nms_host_workspace_bytes(N, A, C, H0, W0, H1, W1));
    CUDA_TRY(cudaSetDevice(device));
    HostStreams &hs = t_streams[device];
    for (int s = 0; s < kSlots; ++s)
        if (!hs.s[s]) CUDA_TRY(cudaStreamCreateWithFlags(&hs.s[s], cudaStreamNonBlocking));
    auto d_h0 = [&](int s) { return (float *)(base + (size_t)s * hp.slot_bytes); };
    auto d_h1 = [&](int s) { return (float *)(base + (size_t)s * hp.slot_bytes + hp.b_h0); };
    auto d_out = [&](int s) { return (float *)(base + (size_t)s * hp.slot_bytes + hp.b_h0 + hp.b_h1); };
    auto d_cnt = [&](int s) { return (int *)(base + (size_t)s * hp.slot_bytes + hp.b_h0 + hp.b_h1 + hp.b_out); };
    auto d_ws = [&](int s) { return (void *)(base + (size_t)s * hp.slot_bytes + hp.b_h0 + hp.b_h1 + hp.b_out + hp.b_cnt); };
    const size_t K = hp.K, per0 = hp.per0, per1 = hp.per1;
    size_t d2h = 0;
    // one chunk: H2D -> kernel -> D2H(counts) on the slot's stream
    auto push_chunk = [&](int b0, int n, int slot) -> int {
        cudaStream_t st = hs.s[slot];
        CUDA_TRY(cudaMemcpyAsync(d_h0(slot), head0 + per0 * b0, per0 * n * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_h1(slot), head1 + per1 * b0, per1 * n * 4, cudaMemcpyHostToDevice, st));
        const int rc = hp.large ? b200yolo_decode_nms_large(d_h0(slot), d_h1(slot), n, A, C, H0, W0, H1, W1, anchor_wh, conf_thr,
                                                            iou_thr, d_out(slot), d_cnt(slot), nullptr, d_ws(slot), hp.b_ws, (void *)st)
                                : b200yolo_decode_nms(d_h0(slot), d_h1(slot), n, A, C, H0, W0, H1, W1, anchor_wh, conf_thr,
                                                      iou_thr, d_out(slot), d_cnt(slot), nullptr, (void *)st);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(out_count + b0, d_cnt(slot), (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        return 0;
    };
    // ... and, once its counts are on the host, the rows: the first max-count rows of every image in one strided copy
    auto pull_rows = [&](int b0, int n, int slot) -> int {
        cudaStream_t st = hs.s[slot];
        CUDA_TRY(cudaStreamSynchronize(st));
        int mx = 0;
        for (int i = 0; i < n; ++i) mx = out_count[b0 + i] > mx ? out_count[b0 + i] : mx;
        if (mx > (int)K) mx = (int)K;
        d2h += (size_t)n * 4;
        if (mx > 0) {
            CUDA_TRY(cudaMemcpy2DAsync(out + K * 7 * b0, K * 28, d_out(slot), K * 28, (size_t)mx * 28, (size_t)n,
                                       cudaMemcpyDeviceToHost, st));
            d2h += (size_t)mx * 28 * n;
        }
        return 0;
    };
    int rc = 0, slot = 0, prev_b0 = -1, prev_n = 0, prev_slot = 0;
    for (int b0 = 0; b0 < N && rc == 0; b0 += hp.chunk, slot = (slot + 1) % kSlots) {
        const int n = (N - b0 < hp.chunk) ? (N - b0) : hp.chunk;
        // the slot's previous user (three chunks back) must have delivered its rows before its buffers are reused
        if (b0 >= kSlots * hp.chunk) CUDA_TRY(cudaStreamSynchronize(hs.s[slot]));
        rc = push_chunk(b0, n, slot);
        if (rc == 0 && prev_b0 >= 0) rc = pull_rows(prev_b0, prev_n, prev_slot);
        prev_b0 = b0; prev_n = n; prev_slot = slot;
    }
    if (rc == 0 && prev_b0 >= 0) rc = pull_rows(prev_b0, prev_n, prev_slot);
    // (on an error the chunks already queued still read / write the caller's buffers: let them finish either way)
    for (int s = 0; s < kSlots; ++s) {
        cudaError_t e = cudaStreamSynchronize(hs.s[s]);
        if (rc == 0 && e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    t_last_d2h = d2h;
    return rc;
}

extern "C" int b200yolo_decode_nms_host(const float *head0, const float *head1, int N, int A, int C, int H0, int W0,
                                        int H1, int W1, const float *anchor_wh, float conf_thr, double iou_thr,
                                        float *out, int *out_count, int device) {
    if (device < 0 || device >= 16) return fail(B200YOLO_EINVAL, "decode_nms_host: bad device");
    if (N <= 0) return N == 0 ? 0 : fail(B200YOLO_EINVAL, "decode_nms_host: bad argument");
    const size_t need = b200yolo_decode_nms_host_workspace_bytes(N, A, C, H0, W0, H1, W1);
    if (need == 0) return fail(B200YOLO_EINVAL, "decode_nms_host: bad shape");
    CUDA_TRY(cudaSetDevice(device));
    OwnedWs &w = t_owned[device];
    if (w.bytes < need) {   // grow-only, one per (thread, device)
        if (w.ptr) CUDA_TRY(cudaFree(w.ptr));
        w.ptr = nullptr; w.bytes = 0;
        CUDA_TRY(cudaMalloc(&w.ptr, need));
        w.bytes = need;
    }
    return b200yolo_decode_nms_host_ws(head0, head1, N, A, C, H0, W0, H1, W1, anchor_wh, conf_thr, iou_thr, out, out_count, w.ptr,
                                       w.bytes, device);
}
