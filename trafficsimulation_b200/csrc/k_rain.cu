// k_rain.cu -- the rain map of a tick: RainAgent.step (agents/rain.py:61-72) rasterises every cloud as the disc
// {(cx + dx, cy + dy) : dx^2 + dy^2 <= r^2} around (int(x), int(y)), clipped to the grid (cells outside have no CellAgent),
// and RainManager.step (:154-185) clears the cells that rained last tick and sets the union of this tick's discs.
// Cloud motion is float arithmetic on the host (x += dx per tick, :59-60); the discs (cx, cy, r) are the tape.
#include "common.cuh"

namespace tsim {

// one CTA per disc and row chunk: rows cy - r .. cy + r, each a horizontal span found from the integer circle equation
__global__ void __launch_bounds__(128) rain_discs_kernel(int W, int y0, int LH, const int32_t *__restrict__ discs, int n, uint8_t value, uint8_t *rain) {
    const int d = blockIdx.y;
    if (d >= n) return;
    const int cx = discs[3 * d], cy = discs[3 * d + 1], r = discs[3 * d + 2];
    if (r < 0) return;
    for (int dy = -r + (int)blockIdx.x; dy <= r; dy += gridDim.x) {
        const int y = cy + dy - y0;   // window-local row
        if (y < 0 || y >= LH) continue;
        int half = 0;                 // largest dx with dx^2 + dy^2 <= r^2 (integer arithmetic, like the reference's offsets list)
        const long long lim = (long long)r * r - (long long)dy * dy;
        while ((long long)(half + 1) * (half + 1) <= lim) half++;
        const int xa = max(cx - half, 0), xb = min(cx + half, W - 1);
        for (int x = xa + threadIdx.x; x <= xb; x += blockDim.x) rain[(size_t)y * W + x] = value;
    }
}

}  // namespace tsim

using namespace tsim;

extern "C" tsim_status tsim_rain_discs(const tsim_cfg *cfg, const int32_t *discs, int32_t n_discs, int32_t value, uint8_t *rain_map, void *stream) {
    tsim_status st = check_cfg(cfg);
    if (st != TSIM_OK) return st;
    if (n_discs < 0 || (n_discs > 0 && !discs) || !rain_map || (value != 0 && value != 1)) { set_error("tsim_rain_discs: bad arguments"); return TSIM_ERR_CONFIG; }
    if (n_discs == 0) return TSIM_OK;
    rain_discs_kernel<<<dim3(64, n_discs), 128, 0, (cudaStream_t)stream>>>(cfg->width, cfg->win_y0, cfg->win_rows, discs, n_discs, (uint8_t)value, rain_map);
    TSIM_LAUNCH_CHECK();
    return TSIM_OK;
}
