// render_cuda_plugin.cpp — the four plugin callbacks + POPULATE_PLUGIN_FUNCTIONS
// (util/render_plugin.h:55-63), mirroring backends/embree/render_embree_plugin.cpp:7-27.
#include <SDL.h>
#include "imgui.h"
#include "render_cuda.h"
#include "render_plugin.h"

#ifdef CRT_CUDA_HEADLESS
// Headless builds (no SDL2 / OpenGL on the machine): a display that shows nothing.
struct NullDisplay : Display {
    std::string gpu_brand() override { return "NVIDIA B200 (headless)"; }
    std::string name() override { return "null"; }
    void resize(const int, const int) override {}
    void new_frame() override {}
    void display(RenderBackend *) override {}
};
#else
#include "display/gldisplay.h"
#endif

uint32_t get_sdl_window_flags()
{
    return SDL_WINDOW_OPENGL;
}

void set_imgui_context(ImGuiContext *context)
{
    ImGui::SetCurrentContext(context);
}

std::unique_ptr<Display> make_display(SDL_Window *window)
{
#ifdef CRT_CUDA_HEADLESS
    (void)window;
    return std::make_unique<NullDisplay>();
#else
    return std::make_unique<GLDisplay>(window);
#endif
}

std::unique_ptr<RenderBackend> make_renderer(Display *)
{
    return std::make_unique<RenderCUDA>();
}

POPULATE_PLUGIN_FUNCTIONS(get_sdl_window_flags, set_imgui_context, make_display, make_renderer)
