// Host-side checks of the __host__ __device__ vector call surface (planet_call_surface.cuh)
// against the arithmetic vec3.h / math.h define.  Built with nvcc, runs without a GPU.
#include <cassert>
#include <cmath>
#include <cstdio>

#include "../planet_b200/csrc/planet_call_surface.cuh"

int main()
{
    Vec3d a = V3d(1.0, -2.0, 3.5), b = V3d(-0.25, 4.0, 2.0);
    Vec3d s = a + b, d = a - b, m = a * 2.0, m2 = 2.0 * a, q = a / 4.0, n = -a;
    assert(s.x == 0.75 && s.y == 2.0 && s.z == 5.5);
    assert(d.x == 1.25 && d.y == -6.0 && d.z == 1.5);
    assert(m.x == 2.0 && m2.z == 7.0 && q.y == -0.5 && n.z == -3.5);
    assert(Dot(a, b) == 1.0 * -0.25 + -2.0 * 4.0 + 3.5 * 2.0);                 // vec3.h:46 order
    assert(LengthSq(a) == 1.0 + 4.0 + 12.25 && Length(a) == std::sqrt(17.25));
    Vec3d u = Normalize(a);                                                     // vec3.h:49: v / Length(v)
    assert(u.x == 1.0 / std::sqrt(17.25) && u.y == -2.0 / std::sqrt(17.25));
    Vec3d c = Cross(a, b);                                                      // vec3.h:59-66
    assert(c.x == -2.0 * 2.0 - 3.5 * 4.0 && c.y == 3.5 * -0.25 - 1.0 * 2.0 && c.z == 1.0 * 4.0 - -2.0 * -0.25);
    Vec3d z = SafeNormalize(V3d(1e-3, 0.0, 0.0));                               // vec3.h:51-57: len2 < epsilon -> zero
    assert(z.x == 0.0 && z.y == 0.0 && z.z == 0.0);
    Vec3 f = V3(a), g = V3(2.0f);
    assert(f.x == 1.0f && f.z == 3.5f && g.y == 2.0f && V3d(3.0).z == 3.0);
    Vec3d e0 = V3d(1.0, 0.0, 0.0), e1 = V3d(0.0, 1.0, 0.0);
    Vec3d h = Slerp(e0, e1, 0.5);                                               // vec3.h:68-72
    assert(std::fabs(h.x - std::sqrt(0.5)) < 1e-7 && std::fabs(h.y - std::sqrt(0.5)) < 1e-7 && h.z == 0.0);
    static_assert(sizeof(Vec3d) == 24 && sizeof(Vec3) == 12, "layout of math.h:42-45");
    printf("call surface ok\n");
    return 0;
}
