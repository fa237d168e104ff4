/* abi_c99.c -- include/sbx.h is a plain C header: this file is compiled with gcc -std=c99 -pedantic -Wall -Werror and
 * linked against libsbx.so; it exercises the entry points that need no device.
 * usage: abi_c99            prints one JSON line; exit code 0 = every check held */
#include <stdio.h>
#include <string.h>

#include "../../include/sbx.h"

#define CHECK(cond) do { if (!(cond)) { printf("{\"failed\": \"%s\", \"line\": %d}\n", #cond, __LINE__); return 1; } } while (0)

int main(void) {
    sbx_params p;
    sbx_shard sh;
    sbx_ctx* ctx = NULL;
    unsigned char hdr[148];
    int rows = 0, part, st;

    CHECK(sbx_default_params(&p, 1920, 1080) == SBX_OK);
    CHECK(p.width == 1920 && p.height == 1080 && p.cld_march_steps == 100 && p.illum_march_steps == 6);
    CHECK(p.sun_dir[2] == -1.0f && p.cld_thick == 125.0f && p.fog_falloff == 0.5f);
    CHECK(sbx_default_params(NULL, 8, 8) == SBX_ERR_INVALID);

    sh.stripe_rows = 4; sh.n_parts = 8;
    for (part = 0; part < 8; ++part) { sh.part = part; rows += sbx_shard_rows(&sh, 1080); }
    CHECK(rows == 1080);
    sh.part = 8;
    CHECK(sbx_shard_rows(&sh, 1080) == SBX_ERR_INVALID);

    CHECK(sbx_dds_volume_header(128, hdr, (int)sizeof hdr) == 148 && memcmp(hdr, "DDS ", 4) == 0);
    CHECK(strcmp(sbx_strerror(SBX_OK), "ok") == 0 && strlen(sbx_strerror(SBX_ERR_NO_DEVICE)) > 0);
    CHECK(strstr(sbx_version(), "sm_100a") != NULL);

    /* no CPU path: without a driver / device the context cannot be created, and every render entry refuses a NULL one */
    st = sbx_create(0, &ctx);
    if (st == SBX_OK) {
        sbx_destroy(ctx);                       /* a GPU box: fine */
    } else {
        CHECK(st == SBX_ERR_NO_DEVICE || st == SBX_ERR_CUDA);
        CHECK(ctx == NULL);
        CHECK(strlen(sbx_last_error(NULL)) > 0);
    }
    CHECK(sbx_render_host(NULL, &p, NULL, (float*)hdr) == SBX_ERR_INVALID);
    CHECK(sbx_render_sequence_host(NULL, &p, NULL, &p.u_time, 1, (float*)hdr) == SBX_ERR_INVALID);
    printf("{\"ok\": true, \"create_status\": %d}\n", st);
    return 0;
}
