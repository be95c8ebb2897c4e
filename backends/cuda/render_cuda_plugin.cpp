// render_cuda_plugin.cpp — what ChameleonRT dlopens as libcrt_cuda.so: the function table of
// util/render_plugin.h:23-41, filled through POPULATE_PLUGIN_FUNCTIONS (util/render_plugin.h:55-63). The role of
// backends/embree/render_embree_plugin.cpp for the Embree backend.
#include "render_plugin.h"

#include <SDL.h>

#include "imgui.h"
#include "render_cuda.h"

#ifndef CRT_CUDA_HEADLESS
#include "display/gldisplay.h"
#endif

namespace crt_cuda_plugin {

#ifdef CRT_CUDA_HEADLESS
// Machines without SDL2 / OpenGL (the build sandbox, the GPU box): a display that shows nothing, so that the
// headless twin of main.cpp (oracle/ref_build) can drive the plugin exactly like the application does.
struct NullDisplay final : Display {
    std::string gpu_brand() override { return "NVIDIA B200 (headless)"; }
    std::string name() override { return "null"; }
    void resize(const int, const int) override {}
    void new_frame() override {}
    void display(RenderBackend *) override {}
};
using PluginDisplay = NullDisplay;
#else
using PluginDisplay = GLDisplay;
#endif

struct Hooks {
    // the window the application creates must be able to host this plugin's display
    static uint32_t window_flags() { return SDL_WINDOW_OPENGL; }

    // the application owns the ImGui context; the plugin's own copy of imgui is pointed at it (main.cpp:98)
    static void adopt_imgui(ImGuiContext *ctx) { ImGui::SetCurrentContext(ctx); }

    static std::unique_ptr<Display> display_for(SDL_Window *window)
    {
#ifdef CRT_CUDA_HEADLESS
        (void)window;
        return std::unique_ptr<Display>(new PluginDisplay());
#else
        return std::unique_ptr<Display>(new PluginDisplay(window));
#endif
    }

    // With this plugin's GLDisplay the frame is presented through CUDA-GL interop (render_cuda.cpp: present_native) and
    // `img` is read back only on request, as backends/optix/render_optix_plugin.cpp:22-26 decides for OptiX; with any
    // other display (the headless NullDisplay) `img` is read back every frame.
    static std::unique_ptr<RenderBackend> renderer_for(Display *display)
    {
#ifdef CRT_CUDA_HEADLESS
        (void)display;
        return std::unique_ptr<RenderBackend>(new RenderCUDA(false));
#else
        return std::unique_ptr<RenderBackend>(new RenderCUDA(dynamic_cast<GLDisplay *>(display) != nullptr));
#endif
    }
};

}  // namespace crt_cuda_plugin

POPULATE_PLUGIN_FUNCTIONS(crt_cuda_plugin::Hooks::window_flags, crt_cuda_plugin::Hooks::adopt_imgui,
                          crt_cuda_plugin::Hooks::display_for, crt_cuda_plugin::Hooks::renderer_for)
