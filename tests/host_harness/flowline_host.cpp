// TEST HARNESS ONLY -- never built into or loaded by the product.
// Compiles the __host__ __device__ core of tendrils_b200/csrc/tb_flowline.cuh for the CPU (g++ -ffp-contract=off) so
// that tests/test_flow_line.py can cross-check the kernels' logic against the oracle on a machine without a GPU.
// The loops below do what k_flow_line_vertices / k_flow_line_raster do: vertex stage for every vertex, then every
// pixel runs all triangles in strip order.
#include "../../tendrils_b200/csrc/tb_flowline.cuh"

#include <vector>

extern "C" long long hh_flow_line(const float *uniforms6, int n, const float *position, const float *normal, const float *miter,
                                  const float *previous, const float *time, const float *dt, float *flow, int W, int H) {
    using namespace tb::fl;
    Uniforms U{uniforms6[0], uniforms6[1], uniforms6[2], uniforms6[3], uniforms6[4], uniforms6[5]};
    std::vector<Vertex> verts(n > 0 ? n : 0);
    for (int i = 0; i < n; ++i)
        vertex_stage(U, position[2 * i], position[2 * i + 1], normal[2 * i], normal[2 * i + 1], miter[i], previous[2 * i],
                     previous[2 * i + 1], time[i], dt[i], W, H, verts[i]);
    long long touched = 0;
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px)
            touched += pixel(verts.data(), n, px, py, U.crestShape, flow + 4 * (static_cast<size_t>(py) * W + px)) ? 1 : 0;
    return touched;
}
