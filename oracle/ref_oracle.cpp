/* TEST INFRASTRUCTURE ONLY -- the reference itself, compiled as a CPU oracle.
 *
 * This translation unit #includes the reference's own main.cpp (which pulls in
 * perlin.h, math.h, vec3.h, list.h, timing.h, pp.h, logging.h, render.h) from
 * where it lies under /root/reference, UNMODIFIED, and links the reference's
 * own render.cpp.  No reference source is copied into this repository: the
 * build recipe (oracle/Makefile, target `ref`) passes `-iquote $(REF)` and the
 * headless SDL/GL stubs in oracle/stubs/.  Output goes to oracle/_ref/ only
 * (git-ignored, shipped to the GPU box as a prebuilt .so).
 *
 * What is the reference's code and what is ours in here:
 *   - every number returned by a ref_* function is computed by reference code:
 *     PerlinRandom/PerlinGradient/PerlinNoise3 (perlin.h:38-88), PerlinfBm /
 *     PerlinRidged (main.cpp:689-734), CreateHeightMapGenerator<F>'s
 *     GetHeightAt / GenerateHeightMap (main.cpp:113-158), QuadID helpers
 *     (main.cpp:19-65), InitPlanet's vertex/index build (main.cpp:391-481),
 *     ProcessQuad / RenderPlanet (main.cpp:537-683), and main() itself;
 *   - ours: the recording fake OpenGL below, and `OracleHeight`, a functor that
 *     restates the 6-line body of `Perlin::operator()` (main.cpp:823-833) with
 *     its constants lifted into a config struct.  `Perlin` is a local struct
 *     inside the reference's main() and cannot be named from outside; to pin
 *     the restated functor, ref_run_reference_main() runs the reference's real
 *     main() for one headless frame and captures the height maps its real
 *     local functor produced (oracle/gen_golden.py asserts both agree bit for
 *     bit).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library.  The product never does.
 */
#include <vector>
#include <string>
#include <thread>
#include <unistd.h>

#define main planet_reference_main
#include "main.cpp"
#undef main

/* ------------------------------------------------------------------------- */
/* recording fake OpenGL                                                      */
/* ------------------------------------------------------------------------- */
namespace fakegl {

struct Buffer { GLenum target; std::vector<unsigned char> bytes; };
struct HeightMap { int w, h; std::vector<float> texels; };
struct DrawRecord {
    float P[12], N[12], skirt_size, corners[4], pixel_size[2];
    GLuint texture; int count;
};

static GLuint next_id = 1;
static std::vector<Buffer> buffers;
static std::vector<HeightMap> height_maps;      /* index = texture id order */
static std::vector<GLuint> height_map_ids;
static std::vector<DrawRecord> draws;
static std::vector<std::string> uniform_names;  /* location = index */
static std::vector<std::string> shader_sources; /* one entry per glShaderSource call */
static float uni_P[12], uni_N[12], uni_skirt, uni_corners[4], uni_pixel[2];
static GLuint bound_texture = 0;
static bool recording = true;

static void reset_frame() { height_maps.clear(); height_map_ids.clear(); draws.clear(); }

} // namespace fakegl

GLboolean glewExperimental = 0;
GLenum glewInit() { return GLEW_OK; }
const unsigned char *glewGetErrorString(GLenum) { return (const unsigned char *)"stub"; }

void glActiveTexture(GLenum) {}
void glAttachShader(GLuint, GLuint) {}
void glBindAttribLocation(GLuint, GLuint, const GLchar *) {}
void glBindBuffer(GLenum, GLuint) {}
void glBindTexture(GLenum, GLuint t) { fakegl::bound_texture = t; }
void glBindVertexArray(GLuint) {}
void glBufferData(GLenum target, GLsizeiptr size, const void *data, GLenum)
{
    fakegl::Buffer b; b.target = target;
    b.bytes.assign((const unsigned char *)data, (const unsigned char *)data + size);
    fakegl::buffers.push_back(b);
}
void glClear(GLbitfield) {}
void glCompileShader(GLuint) {}
GLuint glCreateProgram() { return fakegl::next_id++; }
GLuint glCreateShader(GLenum) { return fakegl::next_id++; }
void glCullFace(GLenum) {}
void glDeleteProgram(GLuint) {}
void glDeleteShader(GLuint) {}
void glDeleteTextures(GLsizei, const GLuint *) {}
void glDepthFunc(GLenum) {}
void glDetachShader(GLuint, GLuint) {}
void glDrawArrays(GLenum, GLint, GLsizei) {}
void glDrawElements(GLenum, GLsizei count, GLenum, const void *)
{
    if (!fakegl::recording) return;
    fakegl::DrawRecord d;
    memcpy(d.P, fakegl::uni_P, sizeof d.P);
    memcpy(d.N, fakegl::uni_N, sizeof d.N);
    d.skirt_size = fakegl::uni_skirt;
    memcpy(d.corners, fakegl::uni_corners, sizeof d.corners);
    memcpy(d.pixel_size, fakegl::uni_pixel, sizeof d.pixel_size);
    d.texture = fakegl::bound_texture;
    d.count = count;
    fakegl::draws.push_back(d);
}
void glEnable(GLenum) {}
void glEnableVertexAttribArray(GLuint) {}
void glFrontFace(GLenum) {}
void glGenBuffers(GLsizei n, GLuint *o) { for (int i = 0; i < n; i++) o[i] = fakegl::next_id++; }
void glGenTextures(GLsizei n, GLuint *o) { for (int i = 0; i < n; i++) o[i] = fakegl::next_id++; }
void glGenVertexArrays(GLsizei n, GLuint *o) { for (int i = 0; i < n; i++) o[i] = fakegl::next_id++; }
GLenum glGetError() { return GL_NO_ERROR; }
void glGetProgramInfoLog(GLuint, GLsizei, GLsizei *, GLchar *b) { if (b) b[0] = 0; }
void glGetProgramiv(GLuint, GLenum, GLint *v) { *v = GL_TRUE; }
void glGetShaderInfoLog(GLuint, GLsizei, GLsizei *, GLchar *b) { if (b) b[0] = 0; }
void glGetShaderiv(GLuint, GLenum, GLint *v) { *v = GL_TRUE; }
GLint glGetUniformLocation(GLuint, const GLchar *name)
{
    for (size_t i = 0; i < fakegl::uniform_names.size(); i++)
        if (fakegl::uniform_names[i] == name) return (GLint)i;
    fakegl::uniform_names.push_back(name);
    return (GLint)fakegl::uniform_names.size() - 1;
}
void glLinkProgram(GLuint) {}
void glPolygonMode(GLenum, GLenum) {}
/* the shader text the reference hands to the driver (render.cpp:111): kept so that the tests can
 * state exactly which text the CPU restatement of the GLSL stage was written against */
void glShaderSource(GLuint, GLsizei count, const GLchar *const *strings, const GLint *lengths)
{
    std::string text;
    for (GLsizei i = 0; i < count; i++)
        text += lengths ? std::string(strings[i], (size_t)lengths[i]) : std::string(strings[i]);
    fakegl::shader_sources.push_back(text);
}
void glTexImage2D(GLenum, GLint, GLint, GLsizei w, GLsizei h, GLint, GLenum, GLenum, const void *data)
{
    if (!fakegl::recording) return;
    fakegl::HeightMap m; m.w = w; m.h = h;
    m.texels.assign((const float *)data, (const float *)data + (size_t)w * h);
    fakegl::height_maps.push_back(m);
    fakegl::height_map_ids.push_back(fakegl::bound_texture);
}
void glTexParameteri(GLenum, GLenum, GLint) {}
static const std::string &uname(GLint loc)
{
    static const std::string none;
    return (loc >= 0 && (size_t)loc < fakegl::uniform_names.size()) ? fakegl::uniform_names[loc] : none;
}
void glUniform1fv(GLint loc, GLsizei, const GLfloat *v) { if (uname(loc) == "SkirtSize") fakegl::uni_skirt = v[0]; }
void glUniform1i(GLint, GLint) {}
void glUniform2fv(GLint loc, GLsizei n, const GLfloat *v)
{
    if (uname(loc) == "HeightMap_corners") memcpy(fakegl::uni_corners, v, sizeof(float) * 2 * (n < 2 ? n : 2));
    if (uname(loc) == "HeightMap_pixel_size") memcpy(fakegl::uni_pixel, v, sizeof(float) * 2);
}
void glUniform3fv(GLint loc, GLsizei n, const GLfloat *v)
{
    if (uname(loc) == "P") memcpy(fakegl::uni_P, v, sizeof(float) * 3 * (n < 4 ? n : 4));
    if (uname(loc) == "N") memcpy(fakegl::uni_N, v, sizeof(float) * 3 * (n < 4 ? n : 4));
}
void glUniformMatrix4fv(GLint, GLsizei, GLboolean, const GLfloat *) {}
void glUseProgram(GLuint) {}
void glVertexAttribPointer(GLuint, GLint, GLenum, GLboolean, GLsizei, const void *) {}
void glViewport(GLint, GLint, GLsizei, GLsizei) {}

/* ------------------------------------------------------------------------- */
/* the height functor, constants lifted out (restates main.cpp:823-833)       */
/* ------------------------------------------------------------------------- */
enum { ORACLE_RIDGED = 0, ORACLE_FBM = 1, ORACLE_ZERO = 2 };

struct OracleFunctorConfig
{
    int kind;            /* main.cpp:829 (ridged) / :830 (fBm, commented out) / :835-841 (zero) */
    double lacunarity;   /* main.cpp:829: 2.0f widened to the double parameter */
    float gain;          /* main.cpp:829: 0.55f */
    int fixed_octaves;   /* <= 0: use main.cpp:827, 6 + 12*depth/max_depth */
    double coord_scale;  /* main.cpp:828: 0.00001 */
    float height_scale;  /* main.cpp:831: 8848.0f */
};

static OracleFunctorConfig g_cfg = { ORACLE_RIDGED, 2.0, 0.55f, 0, 0.00001, 8848.0f };

struct OracleHeight
{
    inline float operator()(Vec3d p, int depth, int max_depth)
    {
        if (g_cfg.kind == ORACLE_ZERO) return 0.0f;
        int octaves = (g_cfg.fixed_octaves > 0) ? g_cfg.fixed_octaves
                                                : 6 + 12 * depth / max_depth;
        p *= g_cfg.coord_scale;
        float h = (g_cfg.kind == ORACLE_RIDGED)
            ? PerlinRidged(p.x, p.y, p.z, g_cfg.lacunarity, g_cfg.gain, octaves)
            : PerlinfBm(p.x, p.y, p.z, g_cfg.lacunarity, g_cfg.gain, octaves);
        return h * g_cfg.height_scale;
    }
};

static HeightMapGenerator g_gen = CreateHeightMapGenerator<OracleHeight>();
static Planet g_planet;            /* zero-initialised like main.cpp:846 */
static bool g_planet_ready = false;
static size_t g_vertex_buffer = 0, g_index_buffer = 0;

static bool ensure_planet(double radius)
{
    if (g_planet_ready && g_planet.radius == radius) return true;
    size_t first = fakegl::buffers.size();
    g_planet = Planet{};
    if (!InitPlanet(g_planet, radius, g_gen)) return false;
    g_vertex_buffer = first;           /* main.cpp:479 */
    g_index_buffer = first + 1;        /* main.cpp:480 */
    g_planet_ready = true;
    return true;
}

extern "C" {

/* ---- perlin.h ---- */
void ref_perlin_tables(unsigned char *table256, float *vectors48)
{
    memcpy(table256, perlin_random_table, 256);
    memcpy(vectors48, perlin_vectors, sizeof(float) * 48);
}
int ref_perlin_random(int seed) { return PerlinRandom(seed); }
float ref_perlin_gradient(float x, float y, float z, int ix, int iy, int iz)
{
    return PerlinGradient(x, y, z, ix, iy, iz);
}
float ref_perlin_noise3(double x, double y, double z) { return PerlinNoise3(x, y, z); }
float ref_perlin_fbm(double x, double y, double z, double lac, float gain, int oct)
{
    return PerlinfBm(x, y, z, lac, gain, oct);
}
float ref_perlin_ridged(double x, double y, double z, double lac, float gain, int oct)
{
    return PerlinRidged(x, y, z, lac, gain, oct);
}
void ref_noise3_batch(const double *xyz, long n, float *out)
{
    for (long i = 0; i < n; i++) out[i] = PerlinNoise3(xyz[3*i], xyz[3*i+1], xyz[3*i+2]);
}
void ref_fractal_batch(const double *xyz, long n, int kind, double lac, float gain, int oct, float *out)
{
    for (long i = 0; i < n; i++)
        out[i] = (kind == ORACLE_RIDGED)
            ? PerlinRidged(xyz[3*i], xyz[3*i+1], xyz[3*i+2], lac, gain, oct)
            : PerlinfBm(xyz[3*i], xyz[3*i+1], xyz[3*i+2], lac, gain, oct);
}

/* ---- QuadID (main.cpp:19-65) ---- */
uint64_t ref_make_root_id(uint64_t root) { return MakeRootID(root).value; }
uint64_t ref_make_child_id(uint64_t id, uint64_t child) { return MakeChildID(QuadID{id}, child).value; }
uint64_t ref_get_parent_id(uint64_t id) { return GetParentID(QuadID{id}).value; }
uint64_t ref_get_root(uint64_t id) { return GetRoot(QuadID{id}); }
uint64_t ref_get_depth(uint64_t id) { return GetDepth(QuadID{id}); }
uint64_t ref_get_index(uint64_t id) { return GetIndex(QuadID{id}); }
uint64_t ref_get_child_index(uint64_t id) { return GetChildIndex(QuadID{id}); }
int ref_sizeof_quad() { return (int)sizeof(Quad); }

/* ---- height functor seam (main.cpp:107-158) ---- */
void ref_set_functor(int kind, double lacunarity, float gain, int fixed_octaves,
                     double coord_scale, float height_scale)
{
    g_cfg.kind = kind; g_cfg.lacunarity = lacunarity; g_cfg.gain = gain;
    g_cfg.fixed_octaves = fixed_octaves; g_cfg.coord_scale = coord_scale;
    g_cfg.height_scale = height_scale;
}
float ref_get_height_at(const double *p, int depth, int max_depth)
{
    Vec3d v = V3d(p[0], p[1], p[2]);
    return g_gen.GetHeightAt(v, depth, max_depth);
}
void ref_generate_height_map(float *data, int dim, const void *quad, int max_depth)
{
    g_gen.GenerateHeightMap(data, dim, *(const Quad *)quad, max_depth);
}
/* the reference is single-threaded; GenerateHeightMap is a pure function of
 * (quad, dim, max_depth) and read-only tables, so patches are striped over
 * threads for the all-cores CPU baseline.  nthreads <= 1: the reference as is. */
void ref_generate_height_maps(const void *quads, long nquads, int dim, int max_depth,
                              float *out, int nthreads)
{
    const Quad *q = (const Quad *)quads;
    size_t per = (size_t)dim * dim;
    if (nthreads <= 1) {
        for (long i = 0; i < nquads; i++) g_gen.GenerateHeightMap(out + i * per, dim, q[i], max_depth);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++)
        pool.emplace_back([=]() {
            for (long i = t; i < nquads; i += nthreads)
                g_gen.GenerateHeightMap(out + i * per, dim, q[i], max_depth);
        });
    for (auto &th : pool) th.join();
}

/* ---- InitPlanet: patch vertex grid + strip index buffer (main.cpp:391-501) ---- */
int ref_init_planet(double radius, int *max_lod, float *max_skirt_size)
{
    if (!ensure_planet(radius)) return 0;
    if (max_lod) *max_lod = g_planet.max_lod;
    if (max_skirt_size) *max_skirt_size = g_planet.max_skirt_size;
    return 1;
}
long ref_patch_buffer(int which, void *out, long cap)
{
    if (!g_planet_ready) return -1;
    const std::vector<unsigned char> &b = fakegl::buffers[which ? g_index_buffer : g_vertex_buffer].bytes;
    if (out && cap >= (long)b.size()) memcpy(out, b.data(), b.size());
    return (long)b.size();
}

/* ---- subdivision geometry through the reference's own ProcessQuad ---- */
/* Root quads: RenderPlanet builds them (main.cpp:604-624) and hands each to
 * ProcessQuad with lod = planet.max_lod; with max_lod forced to 0 ProcessQuad
 * appends them unsplit (main.cpp:539-544). */
int ref_root_quads(double radius, void *out6)
{
    if (!ensure_planet(radius)) return 0;
    OracleFunctorConfig saved = g_cfg; g_cfg.kind = ORACLE_ZERO;
    int saved_lod = g_planet.max_lod; g_planet.max_lod = 0;
    fakegl::recording = false;
    CameraInfo cam = {}; cam.rotation = Mat3Identity();
    InitCameraInfo(cam, DegToRad(50.0f), 800.0f / 600.0f, 1.0f, 20000000.0f);
    RenderPlanet(g_planet, cam);
    fakegl::recording = true;
    g_planet.max_lod = saved_lod; g_cfg = saved;
    if (g_planet.quads.num != 6) return 0;
    memcpy(out6, g_planet.quads.data, 6 * sizeof(Quad));
    return 6;
}
/* One split: ProcessQuad(q, cam, lod = 1) with the camera sitting exactly on
 * corner 0 (zero-height functor, so p[0] == q.p[0] and the distance test at
 * main.cpp:566-573 fires); the four children come back with lod 0 and are
 * appended in child order 0..3 (main.cpp:589-592). */
int ref_split_quad(double radius, const void *quad, void *out4)
{
    if (!ensure_planet(radius)) return 0;
    const Quad &q = *(const Quad *)quad;
    OracleFunctorConfig saved = g_cfg; g_cfg.kind = ORACLE_ZERO;
    CameraInfo cam = {}; cam.position = q.p[0];
    ListResize(g_planet.quads, 0);
    ProcessQuad(g_planet, q, cam, 1);
    g_cfg = saved;
    if (g_planet.quads.num != 4) return 0;
    memcpy(out4, g_planet.quads.data, 4 * sizeof(Quad));
    return 4;
}
/* all quads of one root face at uniform depth, in the recursion (= QuadID path) order
 * ProcessQuad would emit them */
static void uniform_rec(double radius, const Quad &q, int levels, Quad *out, long &n)
{
    if (levels == 0) { out[n++] = q; return; }
    Quad kids[4];
    ref_split_quad(radius, &q, kids);
    for (int c = 0; c < 4; c++) uniform_rec(radius, kids[c], levels - 1, out, n);
}
long ref_uniform_quads(double radius, int face, int depth, void *out)
{
    Quad roots[6];
    if (!ref_root_quads(radius, roots)) return 0;
    long n = 0;
    uniform_rec(radius, roots[face], depth, (Quad *)out, n);
    return n;
}

/* ---- one frame of RenderPlanet with our planet + the configurable functor ---- */
static long render_frame_impl(double radius, const double *cam_pos, bool cold_cache);
long ref_render_frame(double radius, const double *cam_pos)
{
    return render_frame_impl(radius, cam_pos, true);    /* cold cache: every leaf generates */
}
/* the next frame of a sequence: the height-map cache (main.cpp:75-102, 191-278), its LRU ticks
 * and render_tick carry over; captured height maps / draws are those of this frame only */
long ref_render_next_frame(double radius, const double *cam_pos)
{
    return render_frame_impl(radius, cam_pos, false);
}
void ref_reset_cache()
{
    g_planet.cache = HeightMapCache{};
    g_planet.render_tick = 0;
}
static long render_frame_impl(double radius, const double *cam_pos, bool cold_cache)
{
    if (!ensure_planet(radius)) return -1;
    fakegl::reset_frame();
    if (cold_cache) g_planet.cache = HeightMapCache{};
    CameraInfo cam = {};
    cam.position = V3d(cam_pos[0], cam_pos[1], cam_pos[2]);
    cam.rotation = Mat3Identity();
    InitCameraInfo(cam, DegToRad(50.0f), 800.0f / 600.0f, 1.0f, 20000000.0f);
    RenderPlanet(g_planet, cam);
    return g_planet.quads.num;
}
long ref_frame_quads(void *out, long cap)
{
    long n = g_planet.quads.num;
    if (out && cap >= n) memcpy(out, g_planet.quads.data, n * sizeof(Quad));
    return n;
}

/* ---- GetHeightMapForQuad (main.cpp:191-278) on an ARBITRARY quad list, as one frame ---- */
/* The loop of main.cpp:652-660 + :682 for a caller-chosen list (duplicates, a quad together with
 * its parent, quads whose cached entry is evicted earlier in the same list ...): per quad, out8 =
 * { GL texture name, corners[4], pixel_size[2], 1 if this lookup generated a map }.  The functor is
 * switched to ConstantZero for the duration: only the bookkeeping is of interest. */
long ref_cache_lookup(double radius, const void *quads, long n, int budget, float *out8)
{
    if (!ensure_planet(radius)) return -1;
    OracleFunctorConfig saved = g_cfg; g_cfg.kind = ORACLE_ZERO;
    const Quad *q = (const Quad *)quads;
    int left = budget;
    for (long i = 0; i < n; i++) {
        fakegl::reset_frame();
        TextureRect r = GetHeightMapForQuad(g_planet, q[i], left);
        float *o = out8 + 8 * i;
        o[0] = (float)r.texture;
        o[1] = r.corners[0].x; o[2] = r.corners[0].y; o[3] = r.corners[1].x; o[4] = r.corners[1].y;
        o[5] = r.pixel_size.x; o[6] = r.pixel_size.y;
        o[7] = fakegl::height_maps.empty() ? 0.0f : 1.0f;
    }
    g_planet.render_tick++;
    g_cfg = saved;
    return g_planet.cache.count;
}

/* ---- the reference's real main(), one headless frame ---- */
int ref_run_reference_main(const char *scratch_dir)
{
    /* main() reads/writes a file called "save" in the cwd (main.cpp:869-888,1118-1138) */
    if (scratch_dir && chdir(scratch_dir) != 0) return -1;
    fakegl::reset_frame();
    return planet_reference_main(0, nullptr);
}

/* ---- what the fake GL recorded during the last frame ---- */
long ref_captured_height_map_count() { return (long)fakegl::height_maps.size(); }
int ref_captured_height_map(long i, float *out, int *w, int *h)
{
    if (i < 0 || i >= (long)fakegl::height_maps.size()) return 0;
    const fakegl::HeightMap &m = fakegl::height_maps[i];
    if (w) *w = m.w; if (h) *h = m.h;
    if (out) memcpy(out, m.texels.data(), m.texels.size() * sizeof(float));
    return 1;
}
long ref_captured_draw_count() { return (long)fakegl::draws.size(); }
/* shader k as handed to glShaderSource by InitPlanet (0 = vertex stage, 1 = fragment stage);
 * returns its length, copies at most cap-1 bytes */
long ref_captured_shader_count() { return (long)fakegl::shader_sources.size(); }
long ref_captured_shader_source(long k, char *buf, long cap)
{
    if (k < 0 || k >= (long)fakegl::shader_sources.size()) return -1;
    const std::string &t = fakegl::shader_sources[(size_t)k];
    if (buf && cap > 0) {
        long n = std::min<long>((long)t.size(), cap - 1);
        memcpy(buf, t.data(), (size_t)n);
        buf[n] = 0;
    }
    return (long)t.size();
}
/* 32 floats per draw: P[12] N[12] skirt corners[4] pixel_size[2] count */
int ref_captured_draw(long i, float *out32)
{
    if (i < 0 || i >= (long)fakegl::draws.size()) return 0;
    const fakegl::DrawRecord &d = fakegl::draws[i];
    memcpy(out32, d.P, 48); memcpy(out32 + 12, d.N, 48);
    out32[24] = d.skirt_size; memcpy(out32 + 25, d.corners, 16); memcpy(out32 + 29, d.pixel_size, 8);
    out32[31] = (float)d.count;
    return 1;
}
/* GL texture name bound at draw i / given to generated height map i (cache identity) */
long ref_captured_draw_texture(long i)
{
    return (i >= 0 && i < (long)fakegl::draws.size()) ? (long)fakegl::draws[i].texture : -1;
}
long ref_captured_height_map_id(long i)
{
    return (i >= 0 && i < (long)fakegl::height_map_ids.size()) ? (long)fakegl::height_map_ids[i] : -1;
}
int ref_cache_count() { return g_planet.cache.count; }
long ref_captured_buffer_count() { return (long)fakegl::buffers.size(); }
long ref_captured_buffer(long i, void *out, long cap)
{
    if (i < 0 || i >= (long)fakegl::buffers.size()) return -1;
    const std::vector<unsigned char> &b = fakegl::buffers[i].bytes;
    if (out && cap >= (long)b.size()) memcpy(out, b.data(), b.size());
    return (long)b.size();
}

/* ---- vec3.h / math.h: the reference's own vector functions on arrays ---- */
/* a, b: n x 3, t: n.  out: n x 33 = Dot, LengthSq, Length, Normalize[3], SafeNormalize[3], Cross[3],
 * Slerp[3], a+b[3], a-b[3], a*t[3], t*a[3], a/t[3], -a[3]  (vec3.h:25-72), once per instantiation. */
} /* extern "C" */
template <class V, class T> static void vec3_ops(const T *a, const T *b, const T *t, long n, T *out)
{
    for (long i = 0; i < n; i++) {
        V A = { a[3*i], a[3*i+1], a[3*i+2] }, B = { b[3*i], b[3*i+1], b[3*i+2] };
        T *o = out + 33 * i;
        V r[10] = { Normalize(A), SafeNormalize(A), Cross(A, B), Slerp(A, B, t[i]), A + B, A - B, A * t[i], t[i] * A, A / t[i], -A };
        o[0] = Dot(A, B); o[1] = LengthSq(A); o[2] = Length(A);
        for (int k = 0; k < 10; k++) { o[3 + 3*k] = r[k].x; o[4 + 3*k] = r[k].y; o[5 + 3*k] = r[k].z; }
    }
}
extern "C" {
void ref_vec3_ops_f32(const float *a, const float *b, const float *t, long n, float *out) { vec3_ops<Vec3, float>(a, b, t, n, out); }
void ref_vec3_ops_f64(const double *a, const double *b, const double *t, long n, double *out) { vec3_ops<Vec3d, double>(a, b, t, n, out); }

} /* extern "C" */


/* ------------------------------------------------------------------------- */
/* the reference's own caller on the GPU library (-DPLANET_REAL_CALLER)        */
/* ------------------------------------------------------------------------- */
/* The "switch unchanged" claim, executed: this same translation unit -- the reference's main.cpp
 * #include'd above, unmodified -- built as a PROGRAM whose HeightMapGenerator is the GPU one
 * (planet_b200/host/planet_host.h, PLANET_HOST_NO_TYPES: the reference's own Vec3d / Quad /
 * HeightMapGenerator are used, nothing is restated), installed exactly where main.cpp:843-847
 * installs CreateHeightMapGenerator<Perlin>():
 *     InitPlanet(planet, radius, hmap_gen)          main.cpp:847 -> :280, :495
 *     RenderPlanet(planet, cam_info)                main.cpp:1087 -> :600
 * so ProcessQuad's GetHeightAt calls (main.cpp:552, 555) and GetHeightMapForQuad's
 * GenerateHeightMap calls (main.cpp:244) go through the two function pointers into
 * libplanet_gpu.so.  What the reference then hands to glTexImage2D (render.cpp:426) is written to
 * stdout: int64 leaf count, the leaf Quads (104 B each), int64 map count, the maps (w*h floats).
 * tests/test_real_caller.py compares them with what the CPU generator produced in the
 * reference's real main() (golden frame_quads / frame_height_maps).
 *   usage: ref_gpu_caller [frames] [cpu|lod]   (further frames fly 20 km east per frame, reusing the cache;
 *                                           "cpu": the same program with the reference's CPU generator
 *                                           installed instead -- the other side of the comparison on a box
 *                                           that has no /root/reference and no golden for later frames;
 *                                           "lod": the second half of INTEGRATION.md's minimum edit as well --
 *                                           the six ProcessQuad calls of RenderPlanet (main.cpp:604-624) replaced
 *                                           by ONE planet_gpu_select_lod, the reference's own per-leaf
 *                                           GetHeightMapForQuad loop (main.cpp:652-660, :682) behind it)
 */
#ifdef PLANET_REAL_CALLER
#define PLANET_HOST_NO_TYPES
#include "../planet_b200/host/planet_host.h"
#include <cuda_runtime_api.h>

/* RenderPlanet with its quadtree recursion on the GPU: leaves from planet_gpu_select_lod into the
 * reference's planet.quads, then the reference's own cache lookups in the reference's order.  (Draw
 * submission, main.cpp:662-679, is presentation and not needed to see what was uploaded.) */
static bool render_planet_gpu_lod(Planet &planet, const CameraInfo &cam, planet_gpu_quad *d_leaves, int64_t capacity)
{
    planet_gpu_params params;
    planet_gpu_default_params(&params);
    params.radius = planet.radius;
    const double cam_pos[3] = { cam.position.x, cam.position.y, cam.position.z };
    int64_t n = 0;
    if (planet_gpu_select_lod(&params, cam_pos, planet.max_lod, d_leaves, capacity, &n, nullptr) != PLANET_OK) {
        fprintf(stderr, "[ERROR] planet_gpu_select_lod: %s\n", planet_gpu_last_error());
        return false;
    }
    ListResize(planet.quads, (int)n);
    if (cudaMemcpy(planet.quads.data, d_leaves, sizeof(Quad) * n, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
    int generations_per_frame = 100;                                   /* main.cpp:652-653 */
    for (int i = 0; i < planet.quads.num; ++i)                         /* main.cpp:655-660 */
        GetHeightMapForQuad(planet, planet.quads.data[i], generations_per_frame);
    planet.render_tick++;                                              /* main.cpp:682 */
    return true;
}

static void write_all(const void *p, size_t n) { if (fwrite(p, 1, n, stdout) != n) exit(4); }

int main(int argc, char **argv)
{
    const int frames = argc > 1 ? atoi(argv[1]) : 1;
    const double radius = 6371000.0;                                   /* main.cpp:821 */
    const bool cpu = argc > 2 && !strcmp(argv[2], "cpu"), lod = argc > 2 && !strcmp(argv[2], "lod");
    planet_gpu_quad *d_leaves = nullptr;
    const int64_t capacity = 1 << 16;
    HeightMapGenerator hmap_gen = cpu ? g_gen                          /* main.cpp:843 as it is */
                                      : CreateGpuHeightMapGenerator(); /* instead of main.cpp:843 */
    if (!hmap_gen.GenerateHeightMap || !hmap_gen.GetHeightAt) return 3;
    static Planet planet;                                              /* zero-initialised, main.cpp:846 */
    if (!InitPlanet(planet, radius, hmap_gen)) return 1;               /* main.cpp:847 */
    for (int f = 0; f < frames; f++) {
        CameraInfo cam_info = {};
        const double a = 20000.0 * f / radius;
        cam_info.position = V3d(sin(a) * (radius + 10.0), 0.0, -cos(a) * (radius + 10.0));   /* frame 0: main.cpp:864 */
        cam_info.rotation = Mat3Identity();
        InitCameraInfo(cam_info, DegToRad(50.0f), 800.0f / 600.0f, 1.0f, 20000000.0f);      /* main.cpp:1072-1075 */
        fakegl::reset_frame();
        if (lod) {
            if (!d_leaves && cudaMalloc((void **)&d_leaves, sizeof(planet_gpu_quad) * capacity) != cudaSuccess) return 5;
            if (!render_planet_gpu_lod(planet, cam_info, d_leaves, capacity)) return 6;
        } else {
            RenderPlanet(planet, cam_info);                            /* main.cpp:1087 */
        }
        int64_t n = planet.quads.num;
        write_all(&n, sizeof n);
        write_all(planet.quads.data, (size_t)n * sizeof(Quad));
        int64_t m = (int64_t)fakegl::height_maps.size();
        write_all(&m, sizeof m);
        for (const fakegl::HeightMap &hm : fakegl::height_maps) write_all(hm.texels.data(), hm.texels.size() * sizeof(float));
    }
    fflush(stdout);
    if (!cpu) planet_gpu_shutdown();
    return 0;
}
#endif
