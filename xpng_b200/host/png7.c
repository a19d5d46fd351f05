/* png7.c — PNG <-> `.7` conversion without libpng (reference tool: 7/seven.c:1-79, which uses libpng's
 * simplified API).  Host-side C on zlib only; SURVEY.md section 8(f) item 1.
 *
 * --to_7 semantics restated from 7/seven.c:39-65:
 *   - 16-bit ("linear") PNGs are rejected (7/seven.c:48, PNG_FORMAT_FLAG_LINEAR);
 *   - everything else is delivered as 8-bit RGB, or RGBA when the file has an alpha channel or a tRNS chunk
 *     (grey replicated, 1/2/4-bit grey scaled to 8 bits, palettes expanded) — what png_image_finish_read
 *     produces for PNG_FORMAT_RGB / PNG_FORMAT_RGBA on sRGB-encoded files;
 *   - normalize_RGBA (7/seven.c:4-37): alpha == 0 pixels lose their colour, an all-opaque alpha plane is dropped;
 *   - store_7.
 * Adam7-interlaced files are de-interlaced pass by pass.  gAMA/iCCP are ignored (the corpus is sRGB, gAMA 0.45455).
 * --to_png writes filter-0 rows through zlib: valid PNG, pixel-identical, not byte-identical to libpng's file. */
#include "seven.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static u32_t be32(const u8_t *p) { return ((u32_t)p[0] << 24) | ((u32_t)p[1] << 16) | ((u32_t)p[2] << 8) | p[3]; }
static void put32(u8_t *p, u32_t v) { p[0] = (u8_t)(v >> 24); p[1] = (u8_t)(v >> 16); p[2] = (u8_t)(v >> 8); p[3] = (u8_t)v; }

static u8_t *read_file(const char *fn, u64_t *size) {
    FILE *f = fopen(fn, "rb");
    if (!f) return NULL;
    if (fseek(f, 0, SEEK_END)) { fclose(f); return NULL; }
    long n = ftell(f);
    if (n < 0 || fseek(f, 0, SEEK_SET)) { fclose(f); return NULL; }
    u8_t *b = malloc((size_t)n + 1);
    if (!b || fread(b, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(b); return NULL; }
    fclose(f);
    *size = (u64_t)n;
    return b;
}

/* 7/seven.c:4-37 */
static _Bool normalize_rgba(xpng_t *d) {
    if (!d->A) return 0;
    const u64_t n = d->s / 4;
    u8_t *p = d->p;
    _Bool translucent = 0, dirty = 0;
    for (u64_t i = 0; i < n; i++) {
        const u8_t a = p[4 * i + 3];
        if (a == 0 && (p[4 * i] | p[4 * i + 1] | p[4 * i + 2])) { dirty = 1; break; }
        if (a != 255) translucent = 1;
    }
    if (dirty) {
        for (u64_t i = 0; i < n; i++) if (p[4 * i + 3] == 0) p[4 * i] = p[4 * i + 1] = p[4 * i + 2] = 0;
        return 0;
    }
    if (translucent) return 0;
    for (u64_t i = 0; i < n; i++) { p[3 * i] = p[4 * i]; p[3 * i + 1] = p[4 * i + 1]; p[3 * i + 2] = p[4 * i + 2]; }
    d->s = 3 * n; d->A = 0;
    return 0;
}

static int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

/* Decodes a PNG into pm (RGB or RGBA, 8 bits).  Returns 1 on failure. */
_Bool png_load(const char *fn, xpng_t *pm) {
    static const u8_t sig[8] = { 137, 80, 78, 71, 13, 10, 26, 10 };
    u64_t fsz = 0;
    u8_t *file = read_file(fn, &fsz);
    _Bool bad = 1;
    u8_t *idat = NULL, *raw = NULL, *out = NULL;
    if (!file || fsz < 8 + 25 || memcmp(file, sig, 8)) goto done;
    u32_t w = 0, h = 0, depth = 0, ctype = 0, interlace = 0, npal = 0, ntrns = 0;
    u8_t pal[256 * 3], trns[256];
    memset(trns, 255, sizeof trns);
    u64_t idat_len = 0, off = 8;
    _Bool have_ihdr = 0, have_trns = 0, end = 0;
    idat = malloc(fsz);
    if (!idat) goto done;
    while (!end && off + 12 <= fsz) {
        const u32_t len = be32(file + off);
        const u8_t *type = file + off + 4, *data = file + off + 8;
        if (off + 12 + (u64_t)len > fsz) goto done;
        if ((u32_t)crc32(crc32(0, type, 4), data, len) != be32(data + len)) goto done;
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) goto done;
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
            if (data[10] || data[11]) goto done;
            have_ihdr = 1;
        } else if (!memcmp(type, "PLTE", 4)) {
            if (len % 3 || len > 768) goto done;
            npal = len / 3; memcpy(pal, data, len);
        } else if (!memcmp(type, "tRNS", 4)) {
            if (len > 256) goto done;
            ntrns = len; memcpy(trns, data, len); have_trns = 1;
        } else if (!memcmp(type, "IDAT", 4)) {
            memcpy(idat + idat_len, data, len); idat_len += len;
        } else if (!memcmp(type, "IEND", 4)) end = 1;
        off += 12 + (u64_t)len;
    }
    if (!have_ihdr || !end || !w || !h || w > (1u << 24) || h > (1u << 24)) goto done;
    if (depth == 16) goto done;                                   /* 7/seven.c:48: linear formats are refused */
    if (interlace > 1) goto done;
    u32_t ch;
    switch (ctype) {
        case 0: ch = 1; if (depth != 1 && depth != 2 && depth != 4 && depth != 8) goto done; break;
        case 2: ch = 3; if (depth != 8) goto done; break;
        case 3: ch = 1; if ((depth != 1 && depth != 2 && depth != 4 && depth != 8) || !npal) goto done; break;
        case 4: ch = 2; if (depth != 8) goto done; break;
        case 6: ch = 4; if (depth != 8) goto done; break;
        default: goto done;
    }
    /* passes: one for a progressive file, seven for Adam7 (PNG specification, section 8.2) */
    static const u32_t XS[7] = { 0, 4, 0, 2, 0, 1, 0 }, YS[7] = { 0, 0, 4, 0, 2, 0, 1 }, DX[7] = { 8, 8, 4, 4, 2, 2, 1 }, DY[7] = { 8, 8, 8, 4, 4, 2, 2 };
    const u32_t npass = interlace ? 7 : 1;
    const u64_t bpp = (ch * depth + 7) / 8;
    u64_t rawsz = 0;
    for (u32_t ps = 0; ps < npass; ps++) {
        const u64_t pw = interlace ? (w > XS[ps] ? (w - XS[ps] + DX[ps] - 1) / DX[ps] : 0) : w;
        const u64_t ph = interlace ? (h > YS[ps] ? (h - YS[ps] + DY[ps] - 1) / DY[ps] : 0) : h;
        if (pw && ph) rawsz += ph * (1 + (pw * ch * depth + 7) / 8);
    }
    raw = malloc(rawsz ? rawsz : 1);
    if (!raw) goto done;
    {
        z_stream z; memset(&z, 0, sizeof z);
        if (inflateInit(&z) != Z_OK) goto done;
        z.next_in = idat; z.avail_in = (uInt)idat_len; z.next_out = raw; z.avail_out = (uInt)rawsz;
        const int rc = inflate(&z, Z_FINISH);
        const _Bool ok = (rc == Z_STREAM_END || rc == Z_OK || rc == Z_BUF_ERROR) && z.total_out == rawsz;
        inflateEnd(&z);
        if (!ok) goto done;
    }
    const _Bool alpha = ctype == 4 || ctype == 6 || have_trns;
    const u32_t oc = alpha ? 4 : 3;
    out = malloc((u64_t)w * h * oc);
    if (!out) goto done;
    const u32_t maxv = (1u << depth) - 1;
    const u32_t key_g = ntrns >= 2 ? (((u32_t)trns[0] << 8) | trns[1]) : 0xFFFFFFFFu;
    const u32_t key_r = ntrns >= 6 ? (((u32_t)trns[0] << 8) | trns[1]) : 0xFFFFFFFFu, key_gg = ntrns >= 6 ? (((u32_t)trns[2] << 8) | trns[3]) : 0,
                key_b = ntrns >= 6 ? (((u32_t)trns[4] << 8) | trns[5]) : 0;
    u8_t *pass_base = raw;
    for (u32_t ps = 0; ps < npass; ps++) {
        const u64_t pw = interlace ? (w > XS[ps] ? (w - XS[ps] + DX[ps] - 1) / DX[ps] : 0) : w;
        const u64_t ph = interlace ? (h > YS[ps] ? (h - YS[ps] + DY[ps] - 1) / DY[ps] : 0) : h;
        if (!pw || !ph) continue;
        const u64_t rowb = (pw * ch * depth + 7) / 8;
        const u32_t xs = interlace ? XS[ps] : 0, ys = interlace ? YS[ps] : 0, dx = interlace ? DX[ps] : 1, dy = interlace ? DY[ps] : 1;
        for (u64_t y = 0; y < ph; y++) {
            /* un-filter in place (PNG specification, filter method 0) */
            u8_t *row = pass_base + y * (rowb + 1) + 1;
            const u8_t *up = y ? row - (rowb + 1) : NULL;
            const u8_t ft = row[-1];
            if (ft > 4) goto done;
            for (u64_t i = 0; i < rowb; i++) {
                const int a = i >= bpp ? row[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
                int pd = 0;
                if (ft == 1) pd = a; else if (ft == 2) pd = b; else if (ft == 3) pd = (a + b) >> 1; else if (ft == 4) pd = paeth(a, b, c);
                row[i] = (u8_t)(row[i] + pd);
            }
            /* expand to RGB / RGBA at the pass's positions */
            for (u64_t x = 0; x < pw; x++) {
                u8_t *o = out + ((u64_t)(ys + y * dy) * w + (xs + x * dx)) * oc;
                u32_t r, g, b, a = 255;
                if (ctype == 0 || ctype == 3) {
                    u32_t v;
                    if (depth == 8) v = row[x];
                    else { const u32_t per = 8 / depth, sh = (per - 1 - (u32_t)(x % per)) * depth; v = (row[x / per] >> sh) & maxv; }
                    if (ctype == 0) { if (have_trns && v == key_g) a = 0; r = g = b = v * 255u / maxv; }
                    else { if (v >= npal) goto done; r = pal[3 * v]; g = pal[3 * v + 1]; b = pal[3 * v + 2]; a = trns[v]; }
                } else if (ctype == 2) {
                    r = row[3 * x]; g = row[3 * x + 1]; b = row[3 * x + 2];
                    if (have_trns && r == key_r && g == key_gg && b == key_b) a = 0;
                } else if (ctype == 4) { r = g = b = row[2 * x]; a = row[2 * x + 1]; }
                else { r = row[4 * x]; g = row[4 * x + 1]; b = row[4 * x + 2]; a = row[4 * x + 3]; }
                o[0] = (u8_t)r; o[1] = (u8_t)g; o[2] = (u8_t)b;
                if (alpha) o[3] = (u8_t)a;
            }
        }
        pass_base += ph * (rowb + 1);
    }
    pm->p = out; out = NULL; pm->w = w; pm->h = h; pm->A = alpha; pm->s = (u64_t)w * h * oc;
    bad = 0;
done:
    free(file); free(idat); free(raw); free(out);
    return bad;
}

static _Bool write_chunk(FILE *f, const char *type, const u8_t *data, u32_t len) {
    u8_t hdr[8], crc[4];
    put32(hdr, len); memcpy(hdr + 4, type, 4);
    put32(crc, (u32_t)crc32(crc32(0, (const u8_t *)type, 4), data, len));
    return fwrite(hdr, 1, 8, f) != 8 || (len && fwrite(data, 1, len, f) != len) || fwrite(crc, 1, 4, f) != 4;
}

/* Writes pm as an 8-bit RGB / RGBA PNG (filter 0 on every row).  Returns 1 on failure. */
_Bool png_store(const xpng_t *pm, const char *fn) {
    static const u8_t sig[8] = { 137, 80, 78, 71, 13, 10, 26, 10 };
    if (!pm || !pm->p || !fn || !pm->w || !pm->h || pm->w > (1u << 24) || pm->h > (1u << 24)) return 1;
    const u32_t oc = 3 + (u32_t)pm->A;
    if (pm->s != pm->w * pm->h * oc) return 1;
    const u64_t rowb = pm->w * oc, rawsz = (rowb + 1) * pm->h;
    u8_t *raw = malloc(rawsz);
    if (!raw) return 1;
    for (u64_t y = 0; y < pm->h; y++) { raw[y * (rowb + 1)] = 0; memcpy(raw + y * (rowb + 1) + 1, pm->p + y * rowb, rowb); }
    uLongf zcap = compressBound((uLong)rawsz);
    u8_t *zbuf = malloc(zcap);
    _Bool bad = !zbuf || compress2(zbuf, &zcap, raw, (uLong)rawsz, 6) != Z_OK;
    free(raw);
    FILE *f = bad ? NULL : fopen(fn, "wb");
    if (!f) { free(zbuf); return 1; }
    u8_t ihdr[13];
    put32(ihdr, (u32_t)pm->w); put32(ihdr + 4, (u32_t)pm->h); ihdr[8] = 8; ihdr[9] = pm->A ? 6 : 2; ihdr[10] = ihdr[11] = ihdr[12] = 0;
    bad = fwrite(sig, 1, 8, f) != 8 || write_chunk(f, "IHDR", ihdr, 13);
    for (uLongf o = 0; !bad && o < zcap; o += 1u << 20)            /* IDAT chunks of 1 MiB */
        bad = write_chunk(f, "IDAT", zbuf + o, (u32_t)((zcap - o) < (1u << 20) ? (zcap - o) : (1u << 20)));
    bad = bad || write_chunk(f, "IEND", NULL, 0);
    bad |= fclose(f) != 0;
    free(zbuf);
    return bad;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Binary PPM (P6), the pixmap format of the reference's older front end (ancestor/gray.c:667-682, fed by
 * `convert x.png x.ppm` in ancestor/test.rb:12) and of the sintel frames BASELINE.md quotes.
 * ppm_load: "P6", then width, height, maxval separated by white space, '#' comments allowed, ONE white-space byte,
 * then w*h*3 samples.  ancestor/gray.c:671-672 reads the two numbers with sscanf and takes the LAST w*h*3 bytes of the
 * file as samples; for every file `convert` writes the two readings agree.  maxval must be 255 (8-bit samples, the only
 * kind the .7 container holds); 16-bit and rescaled files are refused.
 * ppm_store writes exactly the header of ancestor/gray.c:680: "P6\n%lu %lu\n255\n". */
static int ppm_number(const u8_t *d, u64_t n, u64_t *pos, u64_t *val) {
    u64_t i = *pos, v = 0, digits = 0;
    for (;;) {                                   /* white space and comments before a number */
        if (i >= n) return 1;
        if (d[i] == '#') { while (i < n && d[i] != '\n' && d[i] != '\r') i++; continue; }
        if (d[i] == ' ' || d[i] == '\t' || d[i] == '\n' || d[i] == '\r' || d[i] == '\v' || d[i] == '\f') { i++; continue; }
        break;
    }
    while (i < n && d[i] >= '0' && d[i] <= '9') {
        v = v * 10 + (u64_t)(d[i] - '0');
        if (v > (1ull << 32)) return 1;
        i++; digits++;
    }
    if (!digits) return 1;
    *pos = i; *val = v;
    return 0;
}

_Bool ppm_load(const char *fn, xpng_t *pm) {
    u64_t n = 0, pos = 2, w = 0, h = 0, maxval = 0;
    _Bool bad = 1;
    u8_t *d;
    if (!fn || !pm) return 1;
    d = read_file(fn, &n);
    if (!d) return 1;
    if (n < 11 || d[0] != 'P' || d[1] != '6') goto done;
    if (ppm_number(d, n, &pos, &w) || ppm_number(d, n, &pos, &h) || ppm_number(d, n, &pos, &maxval)) goto done;
    if (maxval != 255 || !w || !h || w > (1u << 24) || h > (1u << 24)) goto done;
    if (pos >= n || !(d[pos] == ' ' || d[pos] == '\t' || d[pos] == '\n' || d[pos] == '\r' || d[pos] == '\v' || d[pos] == '\f')) goto done;
    pos++;
    if (n - pos < w * h * 3) goto done;
    pm->w = w; pm->h = h; pm->A = 0; pm->s = w * h * 3;
    pm->p = malloc(pm->s);
    if (!pm->p) goto done;
    memcpy(pm->p, d + pos, pm->s);
    bad = 0;
done:
    free(d);
    return bad;
}

_Bool ppm_store(const xpng_t *pm, const char *fn) {
    char hdr[100];
    int l;
    FILE *f;
    _Bool bad;
    if (!pm || !fn || !pm->p || !pm->w || !pm->h || pm->A || pm->s != pm->w * pm->h * 3) return 1;   /* P6 has no alpha */
    l = sprintf(hdr, "P6\n%lu %lu\n255\n", (unsigned long)pm->w, (unsigned long)pm->h);
    f = fopen(fn, "wb");
    if (!f) return 1;
    bad = fwrite(hdr, 1, (size_t)l, f) != (size_t)l || fwrite(pm->p, 1, pm->s, f) != pm->s;
    return (_Bool)(fclose(f) != 0 || bad);
}

static _Bool is_ppm(const char *fn) {
    u8_t m[2] = { 0, 0 };
    FILE *f = fopen(fn, "rb");
    if (!f) return 0;
    if (fread(m, 1, 2, f) != 2) m[0] = 0;
    fclose(f);
    return m[0] == 'P' && m[1] == '6';
}

/* 7/seven.c:39-79: ./seven --to_7 example.png example.7   |   ./seven --to_png example.7 example.png
 * Additions that leave the reference's contract alone: --to_7 also takes a P6 file (recognised by its magic), and
 * --to_ppm example.7 example.ppm writes one (RGB only). */
int seven_main(int argc, char **argv) {
    xpng_t pm;
    if (argc == 4 && !strcmp(argv[1], "--to_7")) {
        if (is_ppm(argv[2])) return (int)(ppm_load(argv[2], &pm) || store_7(&pm, argv[3]));
        if (png_load(argv[2], &pm)) return 1;
        return (int)(normalize_rgba(&pm) || store_7(&pm, argv[3]));
    }
    if (argc == 4 && !strcmp(argv[1], "--to_ppm")) {
        if (load_7(argv[2], &pm)) return 1;
        return (int)ppm_store(&pm, argv[3]);
    }
    if (argc == 4 && !strcmp(argv[1], "--to_png")) {
        if (load_7(argv[2], &pm)) return 1;
        return (int)png_store(&pm, argv[3]);
    }
    printf("\n"
           "./seven --to_7   example.png example.7\n"
           "./seven --to_png example.7   example.png\n"
           "\n");
    return 1;
}

#ifdef SEVEN_MAIN
int main(int argc, char **argv) { return seven_main(argc, argv); }
#endif
