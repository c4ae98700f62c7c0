/* Headless stand-in for <GLFW/glfw3.h>; see ../vulkan/vulkan.h.  TEST INFRASTRUCTURE ONLY. */
#ifndef GPUCAD_B200_GLFW_SHIM_H
#define GPUCAD_B200_GLFW_SHIM_H
struct GLFWwindow; struct GLFWmonitor;
#endif
