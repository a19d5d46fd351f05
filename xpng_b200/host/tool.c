/* tool.c — the `tool` command line of the exchange scheme, same contract as the reference's
 * Mirroring_and_Rotating/tool.c:129-142:
 *   tool --(r90|r270|mv|mh|mvh|tl|tr) src.7 res.7
 * The pixmap is turned on the GPU (xpngb_transform); there is no CPU path.  Order of checks as in the reference:
 * argument count, then load_7 (exit 1 without text), then the option name (usage text), then store_7. */
#include "seven.h"
#include "xpng_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv) {
    static const char *const names[] = { "--r90", "--r270", "--mv", "--mh", "--mvh", "--tl", "--tr" };
    if (argc == 4) {
        xpng_t pm;
        int op = -1;
        if (load_7(argv[2], &pm)) return 1;
        for (int i = 0; i < 7; i++) if (!strcmp(argv[1], names[i])) { op = i; break; }
        if (op >= 0) {
            xpngb_ctx *ctx = NULL;
            xpngb_image im = { pm.w, pm.h, 0, pm.A, 0 };
            u8_t *res = malloc(pm.s);
            if (!res) return 1;
            if (xpngb_create(&ctx, 0)) { fprintf(stderr, "tool: no usable CUDA device\n"); return 1; }
            if (xpngb_transform(ctx, op, &im, 1, pm.p, pm.s, 0, res, 0)) {
                fprintf(stderr, "tool: %s\n", xpngb_last_error(ctx));
                return 1;
            }
            xpngb_destroy(ctx);
            pm.p = res; pm.w = im.w; pm.h = im.h;
            return (int)store_7(&pm, argv[3]);
        }
    }
    printf("\n\t./tool --(r90|r270|mv|mh|mvh|tl|tr) src.7 res.7\n\n");
    return 1;
}
