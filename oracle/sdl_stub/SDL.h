/* Headless SDL 1.2 stand-in — TEST INFRASTRUCTURE ONLY (part of oracle/).
 *
 * This is NOT SDL and NOT reference code. It declares just enough of the SDL 1.2
 * surface for the unmodified ttsiodras/renderer sources to compile and run with no
 * display, so that the reference binary can act as the parity oracle and the CPU
 * timing baseline. The framebuffer is a plain 32-bpp XRGB8888
 * buffer (Rmask 0xFF0000, Gmask 0xFF00, Bmask 0xFF, pitch = 4*W); every present
 * (SDL_Flip / SDL_UpdateRect) can dump it to disk, see sdl_stub.cc.
 */
#ifndef ORACLE_SDL_STUB_H
#define ORACLE_SDL_STUB_H

#include <stdint.h>
#include <stddef.h>
/* the real SDL_stdinc.h pulls these in; the reference sources rely on it */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint8_t  Uint8;
typedef uint16_t Uint16;
typedef uint32_t Uint32;
typedef uint64_t Uint64;
typedef int8_t   Sint8;
typedef int16_t  Sint16;
typedef int32_t  Sint32;
typedef int64_t  Sint64;

#define SDL_LIL_ENDIAN 1234
#define SDL_BIG_ENDIAN 4321
#define SDL_BYTEORDER  SDL_LIL_ENDIAN

#define SDL_INIT_VIDEO 0x00000020u

#define SDL_SWSURFACE  0x00000000u
#define SDL_HWSURFACE  0x00000001u
#define SDL_ASYNCBLIT  0x00000004u
#define SDL_HWPALETTE  0x20000000u
#define SDL_DOUBLEBUF  0x40000000u
#define SDL_HWACCEL    0x00000100u

typedef struct SDL_Rect  { Sint16 x, y; Uint16 w, h; } SDL_Rect;
typedef struct SDL_Color { Uint8 r, g, b, unused; } SDL_Color;
typedef struct SDL_Palette { int ncolors; SDL_Color *colors; } SDL_Palette;

typedef struct SDL_PixelFormat {
    SDL_Palette *palette;
    Uint8 BitsPerPixel, BytesPerPixel;
    Uint8 Rloss, Gloss, Bloss, Aloss;
    Uint8 Rshift, Gshift, Bshift, Ashift;
    Uint32 Rmask, Gmask, Bmask, Amask;
    Uint32 colorkey;
    Uint8 alpha;
} SDL_PixelFormat;

typedef struct SDL_Surface {
    Uint32 flags;
    SDL_PixelFormat *format;
    int w, h;
    Uint16 pitch;   /* SDL 1.2 keeps the pitch in 16 bits; 3840*4 = 15360 fits */
    void *pixels;
    SDL_Rect clip_rect;
    int refcount;
} SDL_Surface;

#define SDL_MUSTLOCK(s) 0

/* Events: the headless oracle never produces any. */
enum { SDL_NOEVENT = 0, SDL_KEYDOWN = 2, SDL_KEYUP = 3, SDL_QUIT = 12 };
typedef enum {
    SDLK_UNKNOWN = 0, SDLK_ESCAPE = 27,
    SDLK_0 = 48, SDLK_1, SDLK_2, SDLK_3, SDLK_4, SDLK_5, SDLK_6, SDLK_7, SDLK_8, SDLK_9,
    SDLK_a = 97, SDLK_b, SDLK_c, SDLK_d, SDLK_e, SDLK_f, SDLK_g, SDLK_h, SDLK_i, SDLK_j,
    SDLK_k, SDLK_l, SDLK_m, SDLK_n, SDLK_o, SDLK_p, SDLK_q, SDLK_r, SDLK_s, SDLK_t,
    SDLK_u, SDLK_v, SDLK_w, SDLK_x, SDLK_y, SDLK_z,
    SDLK_UP = 273, SDLK_DOWN, SDLK_RIGHT, SDLK_LEFT,
    SDLK_PAGEUP = 280, SDLK_PAGEDOWN = 281
} SDLKey;
typedef struct SDL_keysym { Uint8 scancode; SDLKey sym; int mod; Uint16 unicode; } SDL_keysym;
typedef struct SDL_KeyboardEvent { Uint8 type, which, state; SDL_keysym keysym; } SDL_KeyboardEvent;
typedef union SDL_Event { Uint8 type; SDL_KeyboardEvent key; } SDL_Event;

int          SDL_Init(Uint32 flags);
void         SDL_Quit(void);
char        *SDL_GetError(void);
SDL_Surface *SDL_SetVideoMode(int w, int h, int bpp, Uint32 flags);
Uint32       SDL_MapRGB(const SDL_PixelFormat *fmt, Uint8 r, Uint8 g, Uint8 b);
Uint32       SDL_MapRGBA(const SDL_PixelFormat *fmt, Uint8 r, Uint8 g, Uint8 b, Uint8 a);
int          SDL_LockSurface(SDL_Surface *s);
void         SDL_UnlockSurface(SDL_Surface *s);
int          SDL_FillRect(SDL_Surface *dst, SDL_Rect *r, Uint32 color);
void         SDL_UpdateRect(SDL_Surface *s, Sint32 x, Sint32 y, Uint32 w, Uint32 h);
int          SDL_Flip(SDL_Surface *s);
void         SDL_WM_SetCaption(const char *title, const char *icon);
Uint32       SDL_GetTicks(void);
void         SDL_Delay(Uint32 ms);
int          SDL_PollEvent(SDL_Event *ev);
int          SDL_EnableUNICODE(int enable);

/* Oracle-only additions (not SDL): deterministic per-pixel random stream used by the
 * ambient-occlusion variant of the reference build (SURVEY.md §8c "AO caveat"). */
void oracle_seed(int x, int y);
int  oracle_rand(void);

#ifdef __cplusplus
}
#endif
#endif
