/* cli.c — the `xpng` command line, same contract as the reference's xpng.c:3-23:
 *   xpng -1|-2|-7 in.7 out.xpng     encode      xpng -3 in.jpg out.xpng   (stub, as in the reference)
 *   xpng -d in.xpng out.7           decode      anything else: usage text, exit status 1
 * Exit status is the library's _Bool (0 ok, 1 failure). */
#include "seven.h"
#include <stdio.h>
#include <string.h>

int main(int argc, char **argv) {
    if (argc == 4 && argv[1][0] == '-' && strlen(argv[1]) == 2) {
        xpng_t pm;
        switch (argv[1][1]) {
            case '1': case '2': case '7':
                return (int)(load_7(argv[2], &pm) || xpng_store((u64_t)(argv[1][1] - '0'), &pm, argv[3]));
            case '3':
                return (int)xpng_from_jpg(argv[2], argv[3]);
            case 'd':
                return (int)(xpng_load(argv[2], &pm) || store_7(&pm, argv[3]));
        }
    }
    printf("\n"
           "encode: ./xpng -[127] example.7    example.xpng\n"
           "        ./xpng -3     example.jpg  example.xpng\n"
           "decode: ./xpng -d     example.xpng example.7\n"
           "\n");
    return 1;
}
