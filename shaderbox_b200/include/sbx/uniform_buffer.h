// uniform_buffer.h -- device binding of the uniform blocks (replaces src/uniform_buffer.h:1-63).
// The values come from the sbx_params block of the launch (include/sbx.h), whose field names are
// the reference's uniform names; the defaults of src/uniform_buffer.h:41-58 are applied on the
// host by sbx_default_params().  `sbx_L` is the first member of sbx_app (sbx_kernel.cuh), so the
// default member initialisers below run after it is set.
#define _begin_ubuffer(name, num) /* uniform block `name` */
#define _end_ubuffer
#define _uniform(type, name, default_value) const type name = sbx_uniform(sbx_L->p.name)
#define _pack(x)

// main block, src/uniform_buffer.h:26-36: shadertoy names on the C++ side
#define u_res iResolution
#define u_time iGlobalTime
#define u_mouse iMouse

#if defined(APP_CLOUDS)
_begin_ubuffer(aux_uniform_buffer_t, b1)
    _uniform(vec3, wind_dir, vec3(0, 0, .2f)) _pack(c0);
    _uniform(vec3, sun_dir, vec3(0, 0, -1)) _pack(c1);
    _uniform(vec3, sun_color, vec3(1.f, .7f, .55f)) _pack(c2);
    _uniform(float, sun_power, (8.f)) _pack(c3.x);
    _uniform(int, cld_march_steps, (100)) _pack(c3.y);
    _uniform(int, illum_march_steps, (6)) _pack(c3.z);
    _uniform(float, sigma_scattering, (.15f)) _pack(c3.w);
    _uniform(float, cld_coverage, (.535f)) _pack(c4.x);
    _uniform(float, cld_thick, (125.f)) _pack(c4.y);
    _uniform(float, atm_radius, (5000.f)) _pack(c4.z);
    _uniform(float, atm_ground_y, (4750.f)) _pack(c4.w);
_end_ubuffer;
#elif defined(APP_SDF_AO)
_begin_ubuffer(aux_uniform_buffer_t, b1)
    _uniform(float, fog_density, (.1f)) _pack(c0.x);
    _uniform(float, fog_falloff, (.5f)) _pack(c0.y);
_end_ubuffer;
#else
_begin_ubuffer(aux_uniform_buffer_t, b1)
_end_ubuffer;
#endif
