/* ORACLE / TEST INFRASTRUCTURE. Declarations only, so that loaders/store_png.h
 * (a never-instantiated template here) parses. The oracle writes raw float
 * framebuffers; PNG encode/decode is done by Python (PIL) in tests. */
#ifndef ORACLE_PNG_SHIM_H
#define ORACLE_PNG_SHIM_H
#include <csetjmp>
#include <cstddef>
typedef struct png_struct_def* png_structp; typedef struct png_info_def* png_infop;
typedef unsigned char png_byte; typedef png_byte* png_bytep; typedef size_t png_size_t; typedef void* png_voidp;
typedef unsigned int png_uint_32;
#define PNG_LIBPNG_VER_STRING "shim"
#define PNG_COLOR_TYPE_RGBA 6
#define PNG_INTERLACE_NONE 0
#define PNG_COMPRESSION_TYPE_BASE 0
#define PNG_FILTER_TYPE_BASE 0
png_voidp png_get_io_ptr(png_structp);
png_structp png_create_write_struct(const char*, void*, void*, void*);
png_infop png_create_info_struct(png_structp);
void png_destroy_read_struct(png_structp*, png_infop*, png_infop*);
void png_destroy_write_struct(png_structp*, png_infop*);
jmp_buf& png_jmpbuf(png_structp);
void png_set_write_fn(png_structp, void*, void (*)(png_structp, png_bytep, png_size_t), void (*)(png_structp));
void png_set_IHDR(png_structp, png_infop, int, int, int, int, int, int, int);
void png_write_info(png_structp, png_infop);
void png_write_row(png_structp, png_bytep);
void png_write_end(png_structp, png_infop);
#endif
