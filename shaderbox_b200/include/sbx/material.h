// material.h -- material table (replaces src/material.h:5-36).  Per-pixel state: `materials`
// is a member of the per-pixel app object.
struct material_t {
    vec3 base_color;
    float metallic;
    float roughness;
    float ior;
    float reflectivity;
    float translucency;
};

#define num_materials 8
#define mat_invalid -1
#define mat_debug 0
_mutable(material_t) materials[num_materials];

// select-by-compare keeps the table in registers (the reference's GLSL branch, :24-33)
SBX_FN material_t get_material(_in(int) index) {
    material_t mat;
    _Pragma("unroll") for (int i = 0; i < num_materials; ++i) {
        if (i == index) {
            mat = materials[i];
            break;
        }
    }
    return mat;
}
