// epilogue.h — closes the translation unit opened by prologue.h and exports, for the tests, thin C
// wrappers that CALL THE REFERENCE'S OWN FUNCTIONS (disney_bsdf.ih, lights.ih, texture2d.ih, lcg_rng.ih,
// util.ih, render_embree.ispc) with the same signatures as the oracle's oracle_kat_* entry points
// (oracle/oracle.cpp), so the two can be compared input by input. Own glue; TEST INFRASTRUCTURE.
int crt_ref_max_path_depth = MAX_PATH_DEPTH;  // util.ih:10

static DisneyMaterial kat_material(const float *m)
{
    DisneyMaterial d;
    d.base_color = make_float3(m[0], m[1], m[2]);
    d.metallic = m[3];
    d.specular = m[4];
    d.roughness = m[5];
    d.specular_tint = m[6];
    d.anisotropy = m[7];
    d.sheen = m[8];
    d.sheen_tint = m[9];
    d.clearcoat = m[10];
    d.clearcoat_gloss = m[11];
    d.ior = m[12];
    d.specular_transmission = m[13];
    return d;
}
}  // namespace ispc

extern "C" {
void refembree_set_max_path_depth(int d) { ispc::crt_ref_max_path_depth = d > 0 ? d : MAX_PATH_DEPTH; }
int refembree_get_max_path_depth() { return ispc::crt_ref_max_path_depth; }

void refispc_kat_rng(unsigned int pixel_id, unsigned int frame_id, unsigned int n, unsigned int *states, float *floats)
{
    ispc::LCGRand rng = ispc::get_rng(pixel_id, frame_id);
    for (unsigned int i = 0; i < n; ++i) {
        floats[i] = ispc::lcg_randomf(rng);
        states[i] = rng.state;
    }
}
// out: f.xyz, pdf
void refispc_kat_disney_eval(const float *mat16, const float *n, const float *w_o, const float *w_i, float *out4)
{
    using namespace ispc;
    const DisneyMaterial mat = kat_material(mat16);
    const float3 nn = make_float3(n[0], n[1], n[2]);
    float3 v_x, v_y;
    ortho_basis(v_x, v_y, nn);
    const float3 wo = make_float3(w_o[0], w_o[1], w_o[2]), wi = make_float3(w_i[0], w_i[1], w_i[2]);
    const float3 f = disney_brdf(mat, nn, wo, wi, v_x, v_y);
    out4[0] = f.x;
    out4[1] = f.y;
    out4[2] = f.z;
    out4[3] = disney_pdf(mat, nn, wo, wi, v_x, v_y);
}
// rng_state in/out; out: f.xyz, pdf, w_i.xyz
void refispc_kat_disney_sample(const float *mat16, const float *n, const float *w_o, unsigned int *rng_state, float *out7)
{
    using namespace ispc;
    const DisneyMaterial mat = kat_material(mat16);
    const float3 nn = make_float3(n[0], n[1], n[2]);
    float3 v_x, v_y;
    ortho_basis(v_x, v_y, nn);
    LCGRand rng;
    rng.state = *rng_state;
    float3 w_i = make_float3(0.f);
    float pdf = 0.f;
    const float3 f = sample_disney_brdf(mat, nn, make_float3(w_o[0], w_o[1], w_o[2]), v_x, v_y, rng, w_i, pdf);
    *rng_state = rng.state;
    out7[0] = f.x;
    out7[1] = f.y;
    out7[2] = f.z;
    out7[3] = pdf;
    out7[4] = w_i.x;
    out7[5] = w_i.y;
    out7[6] = w_i.z;
}
// light20: QuadLight; out: sample_pos.xyz, pdf(p = sample_pos, dir), hit flag, t, hit_pos.xyz
void refispc_kat_light(const float *light20, const float *s2, const float *orig, const float *dir, float *out9)
{
    using namespace ispc;
    QuadLight l;
    std::memcpy(&l, light20, sizeof(l));
    const float3 p = sample_quad_light_position(l, make_float2(s2[0], s2[1]));
    const float3 o = make_float3(orig[0], orig[1], orig[2]), d = make_float3(dir[0], dir[1], dir[2]);
    out9[0] = p.x;
    out9[1] = p.y;
    out9[2] = p.z;
    out9[3] = quad_light_pdf(l, p, o, d);
    float t = 0.f;
    float3 lp = make_float3(0.f);
    const bool hit = quad_intersect(l, o, d, t, lp);
    out9[4] = hit ? 1.f : 0.f;
    out9[5] = hit ? t : 0.f;
    out9[6] = hit ? lp.x : 0.f;
    out9[7] = hit ? lp.y : 0.f;
    out9[8] = hit ? lp.z : 0.f;
}
void refispc_kat_texture(const unsigned char *data, int w, int h, int channels, const float *uv, int n, float *out4n)
{
    using namespace ispc;
    ISPCTexture2D t;
    t.width = w;
    t.height = h;
    t.channels = channels;
    t.data = data;
    for (int i = 0; i < n; ++i) {
        const float4 c = texture(&t, make_float2(uv[2 * i], uv[2 * i + 1]));
        out4n[4 * i] = c.x;
        out4n[4 * i + 1] = c.y;
        out4n[4 * i + 2] = c.z;
        out4n[4 * i + 3] = c.w;
    }
}
void refispc_kat_texture_channel(const unsigned char *data, int w, int h, int channels, const float *uv, int n, int channel,
                                 float *out_n)
{
    using namespace ispc;
    ISPCTexture2D t;
    t.width = w;
    t.height = h;
    t.channels = channels;
    t.data = data;
    for (int i = 0; i < n; ++i) {
        out_n[i] = texture_channel(&t, make_float2(uv[2 * i], uv[2 * i + 1]), channel);
    }
}
void refispc_kat_miss(const float *dirs, int n, float *out3n)
{
    for (int i = 0; i < n; ++i) {
        const ispc::float3 c = ispc::miss_shader(ispc::make_float3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
        out3n[3 * i] = c.x;
        out3n[3 * i + 1] = c.y;
        out3n[3 * i + 2] = c.z;
    }
}
void refispc_kat_ortho_basis(const float *n, float *out6)
{
    ispc::float3 vx, vy;
    ispc::ortho_basis(vx, vy, ispc::make_float3(n[0], n[1], n[2]));
    out6[0] = vx.x;
    out6[1] = vx.y;
    out6[2] = vx.z;
    out6[3] = vy.x;
    out6[4] = vy.y;
    out6[5] = vy.z;
}
}
