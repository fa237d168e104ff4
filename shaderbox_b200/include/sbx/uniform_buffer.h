// uniform_buffer.h -- device binding of the uniforms (what src/uniform_buffer.h:1-63 declares for each language).
// Values come from the sbx_params block of the launch (include/sbx.h): its fields carry the reference's uniform
// names, and the reference's defaults (src/uniform_buffer.h:41-58) are applied on the HOST by sbx_default_params(),
// so nothing here repeats them.  Each uniform becomes a const member of the per-pixel app object, loaded once from
// the kernel's __grid_constant__ parameter; `sbx_L` is the object's first member, so it is set before these run.
#define SBX_UNIFORM(type, name) const type name = sbx_uniform(sbx_L->p.name)

// the shadertoy names of the main block on the C++ side (src/uniform_buffer.h:26-36)
#define u_res iResolution
#define u_time iGlobalTime
#define u_mouse iMouse

#if defined(APP_CLOUDS)       // aux block of the cloud apps (:39-55)
SBX_UNIFORM(vec3, wind_dir);
SBX_UNIFORM(vec3, sun_dir);
SBX_UNIFORM(vec3, sun_color);
SBX_UNIFORM(float, sun_power);
SBX_UNIFORM(int, cld_march_steps);
SBX_UNIFORM(int, illum_march_steps);
SBX_UNIFORM(float, sigma_scattering);
SBX_UNIFORM(float, cld_coverage);
SBX_UNIFORM(float, cld_thick);
SBX_UNIFORM(float, atm_radius);
SBX_UNIFORM(float, atm_ground_y);
#elif defined(APP_SDF_AO)     // (:56-60)
SBX_UNIFORM(float, fog_density);
SBX_UNIFORM(float, fog_falloff);
#endif
