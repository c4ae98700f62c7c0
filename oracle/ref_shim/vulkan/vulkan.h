/* Headless stand-in for <vulkan/vulkan.h>.
 *
 * TEST INFRASTRUCTURE ONLY.  The reference translation unit
 * lattice_files/Gratings.cu reaches imgui_impl_vulkan.h through
 * Gratings.h -> ImguiApp.h, which only needs the Vulkan *type names* to parse;
 * no Vulkan symbol is used by the compute code.  This header declares those
 * names as opaque handles / plain enums so the unmodified reference sources
 * compile in a container without the Vulkan SDK.  It is never part of the
 * product library.
 */
#ifndef GPUCAD_B200_VULKAN_SHIM_H
#define GPUCAD_B200_VULKAN_SHIM_H
#include <stdint.h>
#include <stddef.h>

#define VK_NULL_HANDLE 0
#define SHIM_HANDLE(name) typedef struct name##_T* name;
SHIM_HANDLE(VkInstance) SHIM_HANDLE(VkPhysicalDevice) SHIM_HANDLE(VkDevice)
SHIM_HANDLE(VkQueue) SHIM_HANDLE(VkDescriptorPool) SHIM_HANDLE(VkRenderPass)
SHIM_HANDLE(VkPipelineCache) SHIM_HANDLE(VkCommandBuffer) SHIM_HANDLE(VkPipeline)
SHIM_HANDLE(VkDescriptorSet) SHIM_HANDLE(VkSampler) SHIM_HANDLE(VkImageView)
SHIM_HANDLE(VkCommandPool) SHIM_HANDLE(VkFence) SHIM_HANDLE(VkImage)
SHIM_HANDLE(VkFramebuffer) SHIM_HANDLE(VkSemaphore) SHIM_HANDLE(VkSwapchainKHR)
SHIM_HANDLE(VkSurfaceKHR) SHIM_HANDLE(VkBuffer) SHIM_HANDLE(VkDeviceMemory)
#undef SHIM_HANDLE

typedef uint64_t VkDeviceSize;
typedef uint32_t VkFlags;
typedef int VkResult;
typedef int VkSampleCountFlagBits;
typedef int VkImageLayout;
typedef int VkFormat;
typedef int VkColorSpaceKHR;
typedef int VkPresentModeKHR;
typedef struct VkAllocationCallbacks { void* pUserData; } VkAllocationCallbacks;
typedef struct VkSurfaceFormatKHR { VkFormat format; VkColorSpaceKHR colorSpace; } VkSurfaceFormatKHR;
typedef union VkClearValue { float color[4]; struct { float depth; uint32_t stencil; } depthStencil; } VkClearValue;
typedef struct VkPipelineRenderingCreateInfoKHR { int sType; const void* pNext; } VkPipelineRenderingCreateInfoKHR;
typedef void (*PFN_vkVoidFunction)(void);
#endif
