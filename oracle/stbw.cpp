// ORACLE / TEST INFRASTRUCTURE.  The reference defines STB_IMAGE_WRITE_IMPLEMENTATION
// in src/main.cpp:3, which is not part of the oracle build; image_texture.cpp:213
// still needs stbi_write_png at link time.
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include <stb_image_write.h>
