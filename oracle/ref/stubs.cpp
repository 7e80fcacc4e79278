// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Not shipped, not on the product path.
//
// Headless stand-ins for the two reference translation units that need OpenGL/GLFW and therefore
// cannot be compiled in this image:  src/draw.cpp (immediate-mode GL debug drawing, signatures
// src/draw.h:19-31) and framework/src/window.cpp (GLFW window, framework/include/framework/window.h:20-59).
// On the ray-tracing hot path every draw call is a no-op when enableDebugDraw == false
// (src/draw.cpp:25,59-62,99-103,212-215), so empty bodies preserve behaviour.  The camera only needs
// Window::getAspectRatio (framework/src/window.cpp:379-384) and the four register*Callback no-ops
// (framework/src/trackball.cpp:33-49).
#include "draw.h"
#include <framework/window.h>

bool enableDebugDraw = false;

void drawExampleOfCustomVisualDebug() {}
void drawPlane(const glm::vec3&, const glm::vec3&, const glm::vec3&, const glm::vec3&, const glm::vec3&, float) {}
void drawRay(const Ray&, const glm::vec3&) {}
void drawAABB(const AxisAlignedBox&, DrawMode, const glm::vec3&, float) {}
void debugDrawTriangle(const Vertex&, const Vertex&, const Vertex&) {}
void drawTriangle(const Vertex&, const Vertex&, const Vertex&) {}
void drawMesh(const Mesh&) {}
void drawSphere(const Sphere&) {}
void debugDrawSphere(const Sphere&) {}
void drawSphere(const glm::vec3&, float, const glm::vec3&) {}
void setColor(const glm::vec3&) {}
void drawScene(const Scene&) {}

Window::Window(std::string_view, const glm::ivec2& windowSize, OpenGLVersion glVersion, bool presentable)
    : m_pWindow(nullptr)
    , m_windowSize(windowSize)
    , m_glVersion(glVersion)
    , m_presentable(presentable)
{
}
Window::~Window() {}
void Window::registerMouseButtonCallback(MouseButtonCallback&&) {}
void Window::registerMouseMoveCallback(MouseMoveCallback&&) {}
void Window::registerScrollCallback(ScrollCallback&&) {}
void Window::registerWindowResizeCallback(WindowResizeCallback&&) {}
bool Window::isMouseButtonPressed(int) const { return false; }
glm::vec2 Window::getCursorPos() const { return glm::vec2(0.0f); }
float Window::getAspectRatio() const
{
    if (m_windowSize.x == 0 || m_windowSize.y == 0)
        return 1.0f;
    return float(m_windowSize.x) / float(m_windowSize.y);
}
