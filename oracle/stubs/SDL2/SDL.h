/* TEST INFRASTRUCTURE ONLY -- headless stand-in for <SDL2/SDL.h>.
 *
 * The reference (pgcomp/planet) is an SDL2 + OpenGL program.  This image has
 * neither library, so the oracle build (oracle/Makefile -> oracle/_ref/) puts
 * this directory on the include path and compiles the reference's own
 * main.cpp / render.cpp UNMODIFIED from /root/reference against it.
 * Only the declarations the reference actually names are provided; nothing
 * here computes anything on the terrain path.  The two performance-counter
 * functions are real (std::chrono) so the reference's timing.h works.
 */
#ifndef PLANET_ORACLE_SDL_STUB_H
#define PLANET_ORACLE_SDL_STUB_H

/* the reference relies on SDL.h dragging these in (uint64_t in main.cpp:21,
 * memcpy in list.h:29, uintptr_t in list.h:23) */
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cassert>
#include <chrono>

typedef uint8_t  Uint8;
typedef uint32_t Uint32;
typedef uint64_t Uint64;

struct SDL_Window;
typedef void *SDL_GLContext;

enum { SDL_INIT_VIDEO = 0x20 };
enum { SDL_WINDOWPOS_UNDEFINED = 0x1FFF0000 };
enum { SDL_WINDOW_OPENGL = 2, SDL_WINDOW_RESIZABLE = 0x20 };
enum SDL_GLattr {
    SDL_GL_RED_SIZE, SDL_GL_GREEN_SIZE, SDL_GL_BLUE_SIZE, SDL_GL_DEPTH_SIZE,
    SDL_GL_CONTEXT_MAJOR_VERSION, SDL_GL_CONTEXT_MINOR_VERSION
};
enum { SDL_QUIT = 0x100, SDL_WINDOWEVENT = 0x200, SDL_KEYDOWN = 0x300 };
enum { SDL_WINDOWEVENT_RESIZED = 5 };
enum { KMOD_SHIFT = 3 };

enum SDL_Scancode {
    SDL_SCANCODE_A = 4, SDL_SCANCODE_D = 7, SDL_SCANCODE_F = 9, SDL_SCANCODE_K = 14,
    SDL_SCANCODE_P = 19, SDL_SCANCODE_S = 22, SDL_SCANCODE_T = 23, SDL_SCANCODE_W = 26,
    SDL_SCANCODE_1 = 30, SDL_SCANCODE_2, SDL_SCANCODE_3, SDL_SCANCODE_4, SDL_SCANCODE_5,
    SDL_SCANCODE_6, SDL_SCANCODE_7, SDL_SCANCODE_8, SDL_SCANCODE_9, SDL_SCANCODE_0,
    SDL_SCANCODE_ESCAPE = 41,
    SDL_SCANCODE_F1 = 58, SDL_SCANCODE_F2, SDL_SCANCODE_F3, SDL_SCANCODE_F4, SDL_SCANCODE_F5,
    SDL_SCANCODE_F6, SDL_SCANCODE_F7, SDL_SCANCODE_F8, SDL_SCANCODE_F9, SDL_SCANCODE_F10,
    SDL_SCANCODE_F11, SDL_SCANCODE_F12,
    SDL_SCANCODE_RIGHT = 79, SDL_SCANCODE_LEFT, SDL_SCANCODE_DOWN, SDL_SCANCODE_UP,
    SDL_NUM_SCANCODES = 512
};

struct SDL_Keysym { SDL_Scancode scancode; int sym; unsigned short mod; };
struct SDL_KeyboardEvent { Uint32 type; SDL_Keysym keysym; };
struct SDL_WindowEvent { Uint32 type; Uint8 event; int data1, data2; };
union SDL_Event {
    Uint32 type;
    SDL_WindowEvent window;
    SDL_KeyboardEvent key;
};

/* every call below is a no-op that reports success, except the clock */
inline int  SDL_Init(Uint32) { return 0; }
inline void SDL_Quit() {}
inline const char *SDL_GetError() { return "headless oracle stub"; }
inline int  SDL_GL_SetAttribute(SDL_GLattr, int) { return 0; }
inline SDL_Window *SDL_CreateWindow(const char *, int, int, int, int, Uint32) { return (SDL_Window *)1; }
inline void SDL_DestroyWindow(SDL_Window *) {}
inline SDL_GLContext SDL_GL_CreateContext(SDL_Window *) { return (SDL_GLContext)1; }
inline void SDL_GL_DeleteContext(SDL_GLContext) {}
/* the headless event queue delivers exactly one SDL_QUIT, so the reference's
 * frame loop (main.cpp:900) runs one full frame and exits */
inline int  SDL_PollEvent(SDL_Event *e) {
    static int delivered = 0;
    if (delivered) return 0;
    delivered = 1; e->type = SDL_QUIT; return 1;
}
inline const Uint8 *SDL_GetKeyboardState(int *) { static Uint8 keys[SDL_NUM_SCANCODES]; return keys; }
inline Uint32 SDL_GetTicks() {
    using namespace std::chrono;
    static const steady_clock::time_point t0 = steady_clock::now();
    return (Uint32)duration_cast<milliseconds>(steady_clock::now() - t0).count();
}
inline void SDL_SetWindowTitle(SDL_Window *, const char *) {}
inline void SDL_GL_SwapWindow(SDL_Window *) {}
inline void SDL_Delay(Uint32) {}
inline Uint64 SDL_GetPerformanceFrequency() { return 1000000000ull; }
inline Uint64 SDL_GetPerformanceCounter() {
    using namespace std::chrono;
    return (Uint64)duration_cast<nanoseconds>(steady_clock::now().time_since_epoch()).count();
}

#endif
