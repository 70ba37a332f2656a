// host_pipeline.inl -- b200yolo_decode_nms_host: the reference-facing call with
// HOST buffers (what inference.py:121 / train.py:366 hand to the model is host
// data; bench.py times this as "e2e").  Image chunks are pushed through
// kSlots streams: H2D(head0, head1) -> fused kernel -> D2H(rows, counts), so the
// PCIe copies of neighbouring chunks overlap each other and the kernel.
// Included at the end of b200yolo.cu (uses its helpers).

namespace {

constexpr int kSlots = 3;

struct HostCtx {
    int device = -1;
    cudaStream_t stream[kSlots] = {nullptr, nullptr, nullptr};
    float *d_h0[kSlots] = {nullptr, nullptr, nullptr};
    float *d_h1[kSlots] = {nullptr, nullptr, nullptr};
    float *d_out[kSlots] = {nullptr, nullptr, nullptr};
    int *d_cnt[kSlots] = {nullptr, nullptr, nullptr};
    float *d_ws[kSlots] = {nullptr, nullptr, nullptr};   // records of the large-image path
    size_t cap_h0 = 0, cap_h1 = 0, cap_out = 0, cap_cnt = 0, cap_ws = 0;
};

std::mutex g_host_mu;
HostCtx g_host_ctx[16];

int ensure_ctx(HostCtx &cx, int device, size_t b_h0, size_t b_h1, size_t b_out, size_t b_cnt, size_t b_ws) {
    if (cx.device != device) {
        for (int s = 0; s < kSlots; ++s) CUDA_TRY(cudaStreamCreateWithFlags(&cx.stream[s], cudaStreamNonBlocking));
        cx.device = device;
    }
    auto grow = [&](float **arr, size_t &cap, size_t need) -> int {
        if (need <= cap) return 0;
        for (int s = 0; s < kSlots; ++s) {
            if (arr[s]) CUDA_TRY(cudaFree(arr[s]));
            arr[s] = nullptr;
            CUDA_TRY(cudaMalloc((void **)&arr[s], need));
        }
        cap = need;
        return 0;
    };
    if (int rc = grow(cx.d_h0, cx.cap_h0, b_h0)) return rc;
    if (int rc = grow(cx.d_h1, cx.cap_h1, b_h1)) return rc;
    if (int rc = grow(cx.d_out, cx.cap_out, b_out)) return rc;
    if (int rc = grow((float **)cx.d_cnt, cx.cap_cnt, b_cnt)) return rc;
    if (int rc = grow(cx.d_ws, cx.cap_ws, b_ws)) return rc;
    return 0;
}

}  // namespace

extern "C" int b200yolo_decode_nms_host(const float *head0, const float *head1, int N, int A, int C, int H0, int W0,
                                        int H1, int W1, const float *anchor_wh, float conf_thr, double iou_thr,
                                        float *out, int *out_count, int device) {
    if (!head0 || !head1 || !anchor_wh || !out || !out_count) return fail(B200YOLO_EINVAL, "decode_nms_host: null pointer");
    if (N < 0 || A < 1 || A > kMaxAnchors || C < 1 || H0 < 1 || W0 < 1 || H1 < 1 || W1 < 1 || device < 0 || device >= 16)
        return fail(B200YOLO_EINVAL, "decode_nms_host: bad argument");
    if (N == 0) return 0;
    CUDA_TRY(cudaSetDevice(device));
    const size_t attrs = 5 + (size_t)C;
    const size_t per0 = (size_t)A * attrs * H0 * W0, per1 = (size_t)A * attrs * H1 * W1;  // floats per image
    const size_t K = (size_t)A * H0 * W0 + (size_t)A * H1 * W1;
    // chunk so that ~4 chunks cover the batch, at least 8 images each
    int chunk = (N + 3) / 4;  // 4 chunks: measured best for the 256-image batch (profiles/e2e_sweep.py)
    if (chunk < 8) chunk = 8;
    {
        static const int env_chunk = [] { const char *e = getenv("B200YOLO_HOST_CHUNK"); return e ? atoi(e) : 0; }();
        if (env_chunk > 0) chunk = env_chunk;  // tuning knob: images per pipeline chunk
    }
    if (chunk > N) chunk = N;
    std::lock_guard<std::mutex> lock(g_host_mu);
    HostCtx &cx = g_host_ctx[device];
    // images too large for the fused kernel's shared memory go through b200yolo_decode_nms_large (records in a workspace)
    const bool large = !fits_one_cta((int)K, C, MODE_FUSED);
    const size_t b_ws = large ? b200yolo_decode_nms_large_workspace_bytes(chunk, (int)K) : 0;
    if (int rc = ensure_ctx(cx, device, per0 * chunk * 4, per1 * chunk * 4, K * 7 * chunk * 4, (size_t)chunk * 4, b_ws)) return rc;
    // one chunk: H2D -> kernel -> D2H on the slot's stream
    auto push_chunk = [&](int b0, int n, int slot) -> int {
        cudaStream_t st = cx.stream[slot];
        CUDA_TRY(cudaMemcpyAsync(cx.d_h0[slot], head0 + per0 * b0, per0 * n * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(cx.d_h1[slot], head1 + per1 * b0, per1 * n * 4, cudaMemcpyHostToDevice, st));
        const int rc = large ? b200yolo_decode_nms_large(cx.d_h0[slot], cx.d_h1[slot], n, A, C, H0, W0, H1, W1, anchor_wh, conf_thr,
                                                         iou_thr, cx.d_out[slot], cx.d_cnt[slot], nullptr, cx.d_ws[slot], cx.cap_ws,
                                                         (void *)st)
                             : b200yolo_decode_nms(cx.d_h0[slot], cx.d_h1[slot], n, A, C, H0, W0, H1, W1, anchor_wh, conf_thr,
                                                   iou_thr, cx.d_out[slot], cx.d_cnt[slot], nullptr, (void *)st);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(out_count + b0, cx.d_cnt[slot], (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(out + K * 7 * b0, cx.d_out[slot], K * 7 * n * 4, cudaMemcpyDeviceToHost, st));
        return 0;
    };
    int rc = 0, slot = 0;
    for (int b0 = 0; b0 < N && rc == 0; b0 += chunk, slot = (slot + 1) % kSlots)
        rc = push_chunk(b0, (N - b0 < chunk) ? (N - b0) : chunk, slot);
    if (rc) {  // the chunks already queued still read / write the caller's buffers: let them finish first
        for (int s = 0; s < kSlots; ++s) cudaStreamSynchronize(cx.stream[s]);
        return rc;
    }
    for (int s = 0; s < kSlots; ++s) CUDA_TRY(cudaStreamSynchronize(cx.stream[s]));
    return 0;
}
