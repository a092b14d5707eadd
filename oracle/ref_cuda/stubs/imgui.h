// stand-in for Dear ImGui: VtSimParams::OnGUI (Common.hpp L52-68) is never called headless
#pragma once
typedef int ImGuiSliderFlags;
enum { ImGuiSliderFlags_Logarithmic = 32 };
namespace ImGui {
inline void TextUnformatted(const char*) {}
inline void SameLine() {}
inline void Separator() {}
template <class... A> inline bool SliderInt(A&&...) { return false; }
template <class... A> inline bool SliderFloat(A&&...) { return false; }
template <class... A> inline bool SliderFloat3(A&&...) { return false; }
template <class... A> inline bool Checkbox(A&&...) { return false; }
}
