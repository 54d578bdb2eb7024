// camera.cpp — host plumbing: camera matrix, the benchmark orbit and the light-space matrices.
//
// ~30 flops per frame, computed on the CPU with the same libm calls as the reference and handed to
// the kernels as plain floats (never recomputed on the GPU):
//   Camera::UpdateMV                                reference src/Camera.cc:24-42
//   orbit recurrence of main()                      reference src/renderer.cc:250-252,300-316,485-496
//   light placement                                 reference src/renderer.cc:260-296
//   Light::CalculatePositionInCameraSpace           reference src/Light.cc:162-171
//   Light::CalculateXformFromWorldToLightSpace      reference src/Light.cc:173-192
//   Light::CalculateXformFromCameraToLightSpace     reference src/Light.cc:194-216
#include <cmath>
#include <cstring>

#include "../../../include/b200render.h"
#include "../vec.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {
// The reference evaluates these with *runtime-opaque* or compile-time-constant arguments; libm's
// cosf/sinf must be called on the same float values either way. `volatile` keeps the compiler from
// folding the calls with a different (correctly rounded) algorithm than glibc's.
float cosf_rt(float a) { volatile float v = a; return cosf(v); }
float sinf_rt(float a) { volatile float v = a; return sinf(v); }
const float kMaxi = 1.2f;   // Scene::MaxCoordAfterRescale
}

extern "C" {

void b200r_camera_look_at(const float eye[3], const float lookat[3], float mv[9])
{
    V3 fwd = normalize3(mkv3(lookat[0] - eye[0], lookat[1] - eye[1], lookat[2] - eye[2]));
    V3 zenith = mkv3(0.f, 0.f, 1.f);
    V3 right = normalize3(cross3(fwd, zenith));
    V3 up = normalize3(cross3(right, fwd));
    mv[0] = up.x; mv[1] = up.y; mv[2] = up.z;
    mv[3] = right.x; mv[4] = right.y; mv[5] = right.z;
    mv[6] = fwd.x; mv[7] = fwd.y; mv[8] = fwd.z;
}

void b200r_orbit_init(b200r_orbit* o)
{
    // renderer.cc:250-251: angle1 = 0, angle2 = 0.0f*M_PI/180.f; :300 eye = (maxi*4, 0, 0);
    // :316 dAngle = (coord)(0.3f*M_PI/180.0)
    const float EyeDistanceFactor = 4.0f;
    o->eye[0] = kMaxi * EyeDistanceFactor; o->eye[1] = 0.0f; o->eye[2] = 0.0f;
    o->angle1 = 0.0f;
    o->angle2 = (float)(0.0f * M_PI / 180.f);
    o->d_angle = (float)((0.3f) * M_PI / 180.0);
}

void b200r_orbit_step(b200r_orbit* o, float eye_out[3], float mv_out[9])
{
    // renderer.cc:485-494 (autoRotate branch); all-float arithmetic, float cos/sin overloads
    o->angle1 -= o->d_angle;
    float lookat[3] = {0.f, 0.f, 0.f};
    float distance = sqrtf(o->eye[0] * o->eye[0] + o->eye[1] * o->eye[1] + o->eye[2] * o->eye[2]);
    float c2 = cosf_rt(o->angle2), s2 = sinf_rt(o->angle2);
    float ex = distance * c2 * cosf_rt(o->angle1);
    float ey = distance * c2 * sinf_rt(o->angle1);
    float ez = distance * s2;
    o->eye[0] = ex; o->eye[1] = ey; o->eye[2] = ez;
    eye_out[0] = ex; eye_out[1] = ey; eye_out[2] = ez;
    b200r_camera_look_at(o->eye, lookat, mv_out);
}

void b200r_default_light_pos(int index, float pos[3])
{
    const float LightDistanceFactor = 4.0f;
    if (index == 0) {
        // renderer.cc:252 angle3 = 45.0f*M_PI/180.f ; :277-285
        float angle3 = (float)(45.0f * M_PI / 180.f);
        pos[0] = LightDistanceFactor * kMaxi * cosf_rt(angle3);
        pos[1] = LightDistanceFactor * kMaxi * sinf_rt(angle3);
        pos[2] = LightDistanceFactor * kMaxi;
    } else {
        // renderer.cc:288-296
        pos[0] = LightDistanceFactor * kMaxi;
        pos[1] = -LightDistanceFactor * kMaxi;
        pos[2] = LightDistanceFactor * kMaxi;
    }
}

void b200r_light_in_camera_space(const float lp[3], const float eye[3], const float mv[9], float out[3])
{
    V3 r = mat3_mul(mv, v3_from(lp) - v3_from(eye));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

static void light_axes(const float lp[3], V3& up, V3& right, V3& fwd)
{
    fwd = normalize3(mkv3(-lp[0], -lp[1], -lp[2]));
    right = normalize3(cross3(fwd, mkv3(0.f, 0.f, 1.f)));
    up = normalize3(cross3(right, fwd));
}

void b200r_light_world_to_light(const float lp[3], float out[9])
{
    V3 up, right, fwd; light_axes(lp, up, right, fwd);
    out[0] = up.x; out[1] = up.y; out[2] = up.z;
    out[3] = right.x; out[4] = right.y; out[5] = right.z;
    out[6] = fwd.x; out[7] = fwd.y; out[8] = fwd.z;
}

void b200r_light_camera_to_light(const float lp[3], const float mv[9], float out[9])
{
    V3 up, right, fwd; light_axes(lp, up, right, fwd);
    V3 r1 = mat3_mul(mv, up), r2 = mat3_mul(mv, right), r3 = mat3_mul(mv, fwd);
    out[0] = r1.x; out[1] = r1.y; out[2] = r1.z;
    out[3] = r2.x; out[4] = r2.y; out[5] = r2.z;
    out[6] = r3.x; out[7] = r3.y; out[8] = r3.z;
}

void b200r_frame_defaults(b200r_frame* f, uint32_t mode, uint32_t width, uint32_t height,
                          const float eye[3], const float mv[9], uint32_t n_lights)
{
    memset(f, 0, sizeof *f);
    f->mode = mode == 0 ? (uint32_t)B200R_MODE_RAYTRACE_AA : mode;
    f->width = width; f->height = height;
    memcpy(f->eye, eye, 12); memcpy(f->mv, mv, 36);
    f->n_lights = n_lights < 1 ? 1 : (n_lights > B200R_MAX_LIGHTS ? B200R_MAX_LIGHTS : n_lights);
    for (uint32_t i = 0; i < f->n_lights; i++) {
        b200r_default_light_pos((int)i, f->lights[i].pos);
        // main() only refreshes these for mode >= 5 / >= 7 (renderer.cc:498-507); computing them
        // unconditionally is harmless because lower modes never read them.
        b200r_light_in_camera_space(f->lights[i].pos, eye, mv, f->lights[i].in_camera);
        b200r_light_camera_to_light(f->lights[i].pos, mv, f->lights[i].cam2light);
    }
    f->flags = B200R_F_DEFAULT;
    f->ao_samples = 32;
    f->max_depth = 3;
    f->frame_index = 0;
    f->row_first = 0; f->row_step = 1;
}

}  // extern "C"
