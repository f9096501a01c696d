// pf_math.cuh -- transcendental kernels sized for the FP64 pipe of sm_100a.
//
// The mixture log-likelihood (pymixture.py:86-92) costs K exps and one log per (row, node).  With
// the CUDA math library those are ~35-45 FP64 instructions each, every call carries a divergent
// special-case region, and the kernel ends up ISSUE / latency bound.  These versions are
// branch-free on the hot path, keep <= ~2 ulp accuracy (far inside the 1e-10 budget) with 9 / ~24
// FP64 instructions (exp / log: 9 / 10) and short dependency chains, do exponent surgery and range checks on the integer pipe, and take their
// coefficients from the constant bank (a DFMA source operand, no materialising moves).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pf {

constexpr int kExp2TabBits = 8;
constexpr int kExp2TabSize = 1 << kExp2TabBits;
// 2^(j/256), j = 0..255 (copied into shared memory by the kernels: per-lane indexed loads)
// (global, not __constant__: the kernels stage them into shared memory with per-thread indices, which
// would serialise on the constant cache)
__device__ const double kExp2Tab[kExp2TabSize] = {
    1.0, 1.0027112750502025, 1.0054299011128027, 1.0081558981184175,
    1.0108892860517005, 1.0136300849514894, 1.016378314910953, 1.019133996077738,
    1.0218971486541166, 1.0246677928971357, 1.0274459491187637, 1.030231637686041,
    1.0330248790212284, 1.0358256936019572, 1.0386341019613787, 1.041450124688316,
    1.0442737824274138, 1.0471050958792898, 1.0499440858006872, 1.0527907730046264,
    1.0556451783605572, 1.0585073227945128, 1.061377227289262, 1.0642549128844645,
    1.0671404006768237, 1.0700337118202419, 1.0729348675259756, 1.075843889062791,
    1.0787607977571199, 1.0816856149932152, 1.0846183622133092, 1.0875590609177697,
    1.0905077326652577, 1.0934643990728858, 1.0964290818163769, 1.099401802630222,
    1.102382583307841, 1.1053714457017412, 1.1083684117236787, 1.1113735033448175,
    1.1143867425958924, 1.1174081515673693, 1.1204377524096067, 1.12347556733302,
    1.1265216186082418, 1.129575928566288, 1.1326385195987192, 1.1357094141578055,
    1.1387886347566916, 1.1418762039695616, 1.1449721444318042, 1.148076478840179,
    1.1511892299529827, 1.154310420590216, 1.1574400736337511, 1.1605782120274988,
    1.1637248587775775, 1.1668800369524817, 1.1700437696832502, 1.1732160801636373,
    1.1763969916502812, 1.1795865274628758, 1.182784710984341, 1.1859915656609938,
    1.189207115002721, 1.1924313825831512, 1.1956643920398273, 1.1989061670743806,
    1.202156731452703, 1.2054161090051239, 1.2086843236265816, 1.2119613992768012,
    1.215247359980469, 1.2185422298274085, 1.2218460329727576, 1.2251587936371455,
    1.22848053610687, 1.2318112847340759, 1.2351510639369334, 1.2384998981998165,
    1.241857812073484, 1.245224830175258, 1.2486009771892048, 1.2519862778663162,
    1.255380757024691, 1.2587844395497165, 1.2621973503942507, 1.2656195145788063,
    1.2690509571917332, 1.2724917033894028, 1.275941778396392, 1.2794012075056693,
    1.2828700160787783, 1.2863482295460256, 1.2898358734066657, 1.2933329732290895,
    1.2968395546510096, 1.3003556433796506, 1.3038812651919358, 1.3074164459346773,
    1.3109612115247644, 1.3145155879493546, 1.318079601266064, 1.3216532776031575,
    1.3252366431597413, 1.3288297242059544, 1.3324325470831615, 1.3360451382041458,
    1.339667524053303, 1.3432997311868353, 1.3469417862329458, 1.3505937158920345,
    1.3542555469368927, 1.3579273062129011, 1.3616090206382248, 1.365300717204012,
    1.3690024229745905, 1.3727141650876684, 1.3764359707545302, 1.380167867260238,
    1.383909881963832, 1.387662042298529, 1.3914243757719262, 1.3951969099662003,
    1.3989796725383112, 1.4027726912202048, 1.4065759938190154, 1.4103896082172707,
    1.4142135623730951, 1.4180478843204152, 1.4218926021691656, 1.4257477441054942,
    1.42961333839197, 1.433489413367789, 1.4373759974489824, 1.4412731191286257,
    1.4451808069770467, 1.449099089642035, 1.4530279958490526, 1.4569675544014438,
    1.460917794180647, 1.4648787441464057, 1.4688504333369818, 1.4728328908693675,
    1.4768261459394993, 1.4808302278224719, 1.4848451658727524, 1.488870989524397,
    1.4929077282912648, 1.4969554117672355, 1.5010140696264256, 1.5050837316234065,
    1.5091644275934228, 1.5132561874526098, 1.5173590411982147, 1.5214730189088146,
    1.5255981507445384, 1.529734466947287, 1.533881997840956, 1.5380407738316568,
    1.5422108254079407, 1.5463921831410214, 1.550584877685, 1.5547889397770887,
    1.559004400237837, 1.5632312899713576, 1.567469639965553, 1.5717194812923414,
    1.5759808451078865, 1.5802537626528246, 1.5845382652524937, 1.588834384317164,
    1.593142151342267, 1.597461597908627, 1.6017927556826934, 1.606135656416771,
    1.6104903319492543, 1.6148568142048607, 1.6192351351948637, 1.6236253270173289,
    1.6280274218573478, 1.632441451987275, 1.6368674497669644, 1.6413054476440063,
    1.645755478153965, 1.6502175739206177, 1.6546917676561943, 1.6591780921616162,
    1.6636765803267364, 1.6681872651305825, 1.6727101796415966, 1.6772453570178785,
    1.681792830507429, 1.6863526334483934, 1.6909247992693053, 1.6955093614893326,
    1.7001063537185235, 1.7047158096580513, 1.709337763100463, 1.713972247929926,
    1.718619298122478, 1.723278947746274, 1.7279512309618377, 1.732636182022311,
    1.7373338352737062, 1.7420442251551564, 1.746767386199169, 1.7515033530318782,
    1.7562521603732995, 1.761013843037584, 1.7657884359332727, 1.7705759740635547,
    1.7753764925265212, 1.7801900265154245, 1.785016611318935, 1.789856282321401,
    1.7947090750031072, 1.7995750249405351, 1.804454167806624, 1.809346539371032,
    1.8142521755003989, 1.8191711121586085, 1.8241033854070534, 1.8290490314048973,
    1.8340080864093424, 1.8389805867758937, 1.843966568958626, 1.8489660695104508,
    1.8539791250833855, 1.8590057724288205, 1.864046048397789, 1.8690999899412386,
    1.8741676341103, 1.8792490180565602, 1.8843441790323345, 1.8894531543909392,
    1.8945759815869656, 1.8997126981765553, 1.9048633418176741, 1.9100279502703899,
    1.9152065613971474, 1.9203992131630474, 1.925605943636125, 1.930826790987627,
    1.9360617934922943, 1.9413109895286405, 1.9465744175792332, 1.9518521162309783,
    1.9571441241754002, 1.9624504802089273, 1.9677712232331759, 1.9731063922552343,
    1.978456026387951, 1.9838201648502194, 1.9891988469672663, 1.9945921121709402,
};
// ln2^i / i!, i = 1..4: Taylor coefficients of 2^g - 1 (truncation 3.8e-17 relative on |g| <= 1/512)
// Literal (constexpr) rather than __constant__ arrays: ptxas then feeds them to DFMA as uniform-register / immediate
// operands.  A __constant__ array element is LDC'd into a vector register, and a DFMA with three distinct register
// pairs takes 3 register-file cycles instead of the pipe's 2 (scripts/fp64_probe.py: 42.8 vs 60 DFMA/clk/SM).
constexpr double kExp2C1 = 0.6931471805599453, kExp2C2 = 0.2402265069591007, kExp2C3 = 0.055504108664821576,
                 kExp2C4 = 0.009618129107628477;
// {1/c_j, log c_j} for the 256 mantissa intervals [1 + j/256, 1 + (j+1)/256), c_j the midpoint; entries with
// c_j >= sqrt(2) (j >= 106) describe m/2 instead (log c_j - ln 2), so the reduced argument stays within
// [0.707, 1.414] and n ln2 never cancels against it; the two entries adjacent to 1 are exact (1, 0) / (1/2, 0).
// log c_j was evaluated in extended precision for the ROUNDED 1/c_j (oracle-free: numpy longdouble).
constexpr int kLogTabFirstHalf = 106;
__device__ const double kLogTab[512] = {
    1.0, 0.0,
    0.9941747572815534, 0.0058422756242284025,
    0.9903288201160542, 0.00971824946892134,
    0.9865125240847784, 0.013579258126380866,
    0.982725527831094, 0.01742541671385917,
    0.9789674952198852, 0.021256839025415152,
    0.9752380952380952, 0.025073637552115963,
    0.9715370018975332, 0.02887592350185452,
    0.9678638941398866, 0.03266380681879157,
    0.9642184557438794, 0.03643739620243109,
    0.9606003752345216, 0.04019679912633676,
    0.9570093457943926, 0.04394212185649872,
    0.9534450651769087, 0.04767346946935691,
    0.9499072356215214, 0.051390945869489356,
    0.9463955637707948, 0.05509465380697375,
    0.9429097605893186, 0.058784694894427614,
    0.9394495412844037, 0.06246116962373623,
    0.9360146252285192, 0.06612417738247342,
    0.9326047358834244, 0.06977381647002284,
    0.9292196007259528, 0.07341018411340675,
    0.9258589511754068, 0.07703337648282704,
    0.9225225225225225, 0.08064348870692671,
    0.9192100538599641, 0.0842406148877762,
    0.9159212880143113, 0.08782484811559137,
    0.9126559714795008, 0.09139628048318858,
    0.9094138543516874, 0.09495500310018257,
    0.9061946902654867, 0.09850110610693318,
    0.9029982363315696, 0.10203467868824434,
    0.8998242530755711, 0.1055558090868232,
    0.8966725043782837, 0.10906458461650242,
    0.893542757417103, 0.11256109167523176,
    0.8904347826086957, 0.11604541575784262,
    0.8873483535528596, 0.11951764146859183,
    0.8842832469775475, 0.1229778525334875,
    0.8812392426850258, 0.12642613181240342,
    0.8782161234991424, 0.12986256131098461,
    0.8752136752136752, 0.1332872221923487,
    0.8722316865417377, 0.13670019478858866,
    0.8692699490662139, 0.14010155861207893,
    0.8663282571912013, 0.14349139236659045,
    0.863406408094435, 0.1468697739582177,
    0.8605042016806723, 0.1502367805061219,
    0.8576214405360134, 0.15359248835309428,
    0.8547579298831386, 0.15693697307594154,
    0.8519134775374376, 0.16027030949569976,
    0.8490878938640133, 0.16359257168767763,
    0.8462809917355372, 0.1669038329913337,
    0.8434925864909391, 0.17020416601999047,
    0.8407224958949097, 0.17349364267038928,
    0.8379705400981997, 0.17677233413208748,
    0.835236541598695, 0.18004031089670358,
    0.832520325203252, 0.18329764276701008,
    0.8298217179902755, 0.1865443988658801,
    0.827140549273021, 0.18978064764508826,
    0.8244766505636071, 0.19300645689397097,
    0.8218298555377207, 0.19622189374794538,
    0.8192, 0.19942702469689366,
    0.8165869218500797, 0.20262191559341292,
    0.8139904610492846, 0.2058066316609327,
    0.8114104595879557, 0.20898123750170533,
    0.8088467614533965, 0.2121457971046684,
    0.8062992125984252, 0.21530037385318387,
    0.8037676609105181, 0.21844503053265552,
    0.8012519561815337, 0.221579829338027,
    0.7987519500780031, 0.22470483188116222,
    0.7962674961119751, 0.2278200991981116,
    0.7937984496124031, 0.2309256917562647,
    0.7913446676970634, 0.2340216694613928,
    0.7889060092449923, 0.2371080916645822,
    0.7864823348694316, 0.24018501716906146,
    0.7840735068912711, 0.24325250423692324,
    0.7816793893129771, 0.24631061059574413,
    0.7792998477929984, 0.24935939344510277,
    0.7769347496206374, 0.2523989094629994,
    0.7745839636913767, 0.25542921481217845,
    0.7722473604826546, 0.25845036514635467,
    0.7699248120300752, 0.2614624156163463,
    0.767616191904048, 0.26446542087611596,
    0.7653213751868461, 0.2674594350887206,
    0.7630402384500745, 0.270444511932174,
    0.7607726597325408, 0.27342070460522006,
    0.7585185185185185, 0.2763880658330221,
    0.7562776957163959, 0.27934664787276714,
    0.7540500736377025, 0.28229650251918836,
    0.7518355359765051, 0.28523768111000464,
    0.7496339677891655, 0.28817023453128227,
    0.7474452554744525, 0.29109421322271756,
    0.745269286754003, 0.2940096671828415,
    0.7431059506531205, 0.2969166459741508,
    0.7409551374819102, 0.2998151987281622,
    0.7388167388167388, 0.30270537415039545,
    0.7366906474820144, 0.3055872205252843,
    0.7345767575322812, 0.30846078572101604,
    0.7324749642346209, 0.3113261171943025,
    0.7303851640513552, 0.3141832619950823,
    0.7283072546230441, 0.3170322667711571,
    0.7262411347517731, 0.3198731777727608,
    0.7241867043847242, 0.32270604085706495,
    0.7221438645980254, 0.3255309014926197,
    0.720112517580872, 0.3283478047637331,
    0.7180925666199158, 0.3311567953747882,
    0.7160839160839161, 0.3339579176544999,
    0.7140864714086471, 0.3367512155601126,
    0.7121001390820584, 0.33953673268153894,
    0.710124826629681, 0.3423145122454413,
    0.7081604426002767, 0.3450845971192569,
    0.7062068965517241, -0.34530015074477827,
    0.7042640990371389, -0.34254532806593374,
    0.7023319615912208, -0.3397980735907949,
    0.7004103967168263, -0.3370583458496746,
    0.6984993178717599, -0.3343261037128016,
    0.6965986394557823, -0.33160130638661633,
    0.694708276797829, -0.3288839134101164,
    0.6928281461434371, -0.3261738846512514,
    0.6909581646423751, -0.3234711803033662,
    0.6890982503364738, -0.32077576088169396,
    0.687248322147651, -0.3180875872198935,
    0.6854082998661312, -0.31540662046663576,
    0.6835781041388518, -0.3127328220822336,
    0.681757656458056, -0.31006615383531844,
    0.6799468791500664, -0.30740657779955943,
    0.6781456953642384, -0.3047540563504284,
    0.6763540290620872, -0.30210855216200433,
    0.6745718050065876, -0.29947002820382324,
    0.6727989487516426, -0.2968384477377673,
    0.6710353866317169, -0.2942137743149961,
    0.6692810457516339, -0.2915959717729172,
    0.6675358539765319, -0.28898500423219686,
    0.6657997399219766, -0.28638083609380915,
    0.6640726329442282, -0.28378343203612355,
    0.6623544631306598, -0.2811927570120312,
    0.6606451612903226, -0.2786087762461061,
    0.6589446589446589, -0.2760314552318056,
    0.6572528883183568, -0.27346075972870476,
    0.6555697823303457, -0.27089665575976707,
    0.6538952745849298, -0.26833910960865004,
    0.6522292993630573, -0.2657880878170446,
    0.650571791613723, -0.2632435571820499,
    0.6489226869455006, -0.2607054847535788,
    0.6472819216182049, -0.2581738378317993,
    0.6456494325346784, -0.2556485839646051,
    0.6440251572327044, -0.25312969094512117,
    0.6424090338770388, -0.250617126809238,
    0.6408010012515645, -0.24811085983317846,
    0.6392009987515606, -0.24561085853109388,
    0.6376089663760897, -0.2431170916526914,
    0.6360248447204969, -0.24062952818088976,
    0.6344485749690211, -0.2381481373295043,
    0.6328800988875154, -0.23567288854096133,
    0.6313193588162762, -0.23320375148404018,
    0.6297662976629766, -0.23074069605164252,
    0.6282208588957056, -0.22828369235859045,
    0.6266829865361077, -0.22583271073945013,
    0.6251526251526252, -0.22338772174638366,
    0.6236297198538368, -0.2209486961470248,
    0.6221142162818954, -0.2185156049223832,
    0.6206060606060606, -0.2160884192647721,
    0.619105199516324, -0.21366711057576174,
    0.617611580217129, -0.21125165046415806,
    0.6161251504211793, -0.20884201074400494,
    0.6146458583433373, -0.20643816343261034,
    0.6131736526946108, -0.2040400807485976,
    0.6117084826762246, -0.20164773510997772,
    0.6102502979737783, -0.19926109913224688,
    0.6087990487514863, -0.19688014562650494,
    0.6073546856465006, -0.1945048475975977,
    0.6059171597633136, -0.19213517824227927,
    0.6044864226682408, -0.18977111094739865,
    0.6030624263839811, -0.1874126192881057,
    0.6016451233842538, -0.18505967702607892,
    0.6002344665885111, -0.18271225810777392,
    0.5988304093567252, -0.1803703366626929,
    0.5974329054842473, -0.17803388700167322,
    0.5960419091967404, -0.1757028836151977,
    0.59465737514518, -0.1733773011717223,
    0.593279258400927, -0.17105711451602512,
    0.5919075144508671, -0.16874229866757384,
    0.5905420991926182, -0.16643282881891128,
    0.5891829689298044, -0.16412868033406106,
    0.5878300803673938, -0.16182982874695032,
    0.586483390607102, -0.15953624975985092,
    0.5851428571428572, -0.15724791924183873,
    0.5838084378563284, -0.1549648132272701,
    0.5824800910125142, -0.15268690791427605,
    0.5811577752553916, -0.15041417966327356,
    0.579841449603624, -0.14814660499549315,
    0.5785310734463277, -0.14588416059152362,
    0.5772266065388951, -0.14362682328987356,
    0.5759280089988752, -0.1413745700855486,
    0.574635241301908, -0.13912737812864387,
    0.5733482642777156, -0.13688522472295406,
    0.5720670391061452, -0.13464808732459763,
    0.5707915273132664, -0.13241594354065697,
    0.5695216907675195, -0.1301887711278328,
    0.5682574916759157, -0.1279665479911152,
    0.5669988925802879, -0.12574925218246763,
    0.5657458563535912, -0.12353686189952705,
    0.5644983461962514, -0.1213293554843165,
    0.5632563256325632, -0.1191267114219742,
    0.562019758507135, -0.1169289083394948,
    0.56078860898138, -0.11473592500448461,
    0.5595628415300546, -0.11254774032393178,
    0.5583424209378408, -0.11036433334298829,
    0.5571273122959739, -0.10818568324376607,
    0.5559174809989142, -0.1060117693441463,
    0.5547128927410617, -0.10384257109660089,
    0.5535135135135135, -0.10167806808702791,
    0.552319309600863, -0.09951824003359785,
    0.5511302475780409, -0.09736306678561449,
    0.5499462943071965, -0.095212528322386,
    0.5487674169346195, -0.09306660475210926,
    0.5475935828877005, -0.09092527631076612,
    0.5464247598719317, -0.08878852336103096,
    0.5452609158679447, -0.08665632639119028,
    0.5441020191285866, -0.0845286660140734,
    0.542948038176034, -0.0824055229659957,
    0.5417989417989418, -0.0802868781057104,
    0.5406546990496304, -0.07817271241337477,
    0.5395152792413066, -0.07606300698952516,
    0.5383806519453207, -0.07395774305406276,
    0.5372507869884575, -0.07185690194525102,
    0.5361256544502618, -0.06976046511872296,
    0.5350052246603971, -0.06766841414649893,
    0.5338894681960376, -0.0655807307160149,
    0.5327783558792925, -0.0634973966291607,
    0.5316718587746625, -0.06141839380132751,
    0.5305699481865285, -0.059343704260467124,
    0.5294725956566702, -0.05727331014615884,
    0.5283797729618163, -0.055207193708686784,
    0.5272914521112255, -0.05314533730812815,
    0.526207605344296, -0.051087723413448055,
    0.5251282051282051, -0.04903433460160592,
    0.5240532241555783, -0.04698515355667038,
    0.5229826353421859, -0.04494016306894274,
    0.5219164118246687, -0.04289934603409009,
    0.5208545269582909, -0.04086268545228651,
    0.5197969543147208, -0.03883016442736426,
    0.5187436676798379, -0.036801766165971576,
    0.5176946410515673, -0.034777473976741025,
    0.5166498486377397, -0.032757271269465114,
    0.5156092648539778, -0.0307411415542805,
    0.5145728643216081, -0.028729068440860425,
    0.5135406218655968, -0.026721035637614767,
    0.5125125125125125, -0.024717026950899567,
    0.5114885114885115, -0.022717026284232507,
    0.5104685942173479, -0.020721017637517485,
    0.5094527363184079, -0.018728985106276907,
    0.5084409136047666, -0.016740912880890774,
    0.5074331020812686, -0.014756785245844117,
    0.506429277942631, -0.012776586578981642,
    0.5054294175715696, -0.010800301350769646,
    0.5044334975369458, -0.00882791412356531,
    0.5034414945919371, -0.006859409550893202,
    0.5024533856722276, -0.004894772376728207,
    0.5014691478942214, -0.0029339874347875704,
    0.5, 0.0,
};
// 1/(2i+1), i = 1..10: odd atanh series
static __constant__ double kAtanhC[10] = {1.0 / 3.0,  1.0 / 5.0,  1.0 / 7.0,  1.0 / 9.0,  1.0 / 11.0,
                                   1.0 / 13.0, 1.0 / 15.0, 1.0 / 17.0, 1.0 / 19.0, 1.0 / 21.0};
constexpr double kNegHalfLog2e = -0.72134752044448170368;  // -log2(e)/2
constexpr double kLog2e = 1.44269504088896340736;
constexpr double kRintMagic = 26388279066624.0;            // 1.5 * 2^44: rint(256 z) lands in the low word
constexpr double kLn2Hi = 6.93147180369123816490e-01, kLn2Lo = 1.90821492927058770002e-10;

// exp(-e/2) for e >= 0, BRANCH FREE: 9 FP64 instructions + ~7 integer-pipe instructions.
//   2^z,  z = -e*log2(e)/2 = k + j/256 + g,  |g| <= 1/512:
//   2^z = 2^k * tab[j] * (1 + g*(c1 + g*(c2 + g*(c3 + c4 g))))
// k and j come out of one magic-number add, tab[] is a 256-entry shared-memory table, 2^k goes
// through the exponent field.  The exponent is clamped at -1021 on the integer pipe, so terms
// that would be subnormal come out as ~1e-308 instead (harmless in a sum); a row whose whole sum
// is that small -- or inf / nan -- is re-done on the slow path by the caller (needs_slow_path).
__device__ __forceinline__ double exp_neg_half_fast(double e, const double* tab) {
    const double z = e * kNegHalfLog2e;
    const double t = z + kRintMagic;
    const int ki = max(__double2loint(t), -1021 * kExp2TabSize);  // rint(256 z) = 256 k + j
    const double g = z - (t - kRintMagic);
    const int j = ki & (kExp2TabSize - 1);
    const double tj = tab[j];
    const double q = fma(fma(fma(kExp2C4, g, kExp2C3), g, kExp2C2), g, kExp2C1);
    const double p = fma(q * g, tj, tj);  // tab[j] * (1 + g q)
    return __hiloint2double(__double2hiint(p) + (ki >> kExp2TabBits) * (1 << 20), __double2loint(p));
}

// G independent exps, written step-interleaved so the FP64 pipe always has independent work.
template <int G>
__device__ __forceinline__ void exp_neg_half_fast_batch(const double (&e)[G], double (&out)[G], const double* tab) {
    double g[G], q[G], tj[G];
    int kk[G];
#pragma unroll
    for (int c = 0; c < G; ++c) {
        const double z = e[c] * kNegHalfLog2e;
        const double t = z + kRintMagic;
        const int ki = max(__double2loint(t), -1021 * kExp2TabSize);
        g[c] = z - (t - kRintMagic);
        const int j = ki & (kExp2TabSize - 1);
        tj[c] = tab[j];
        kk[c] = ki >> kExp2TabBits;  // k (arithmetic shift: floor)
    }
#pragma unroll
    for (int c = 0; c < G; ++c) q[c] = fma(kExp2C4, g[c], kExp2C3);
#pragma unroll
    for (int c = 0; c < G; ++c) q[c] = fma(q[c], g[c], kExp2C2);
#pragma unroll
    for (int c = 0; c < G; ++c) q[c] = fma(q[c], g[c], kExp2C1);
#pragma unroll
    for (int c = 0; c < G; ++c) {
        const double p = fma(q[c] * g[c], tj[c], tj[c]);
        out[c] = __hiloint2double(__double2hiint(p) + kk[c] * (1 << 20), __double2loint(p));  // one IMAD
    }
}

// Same, with the table given as a 2 KB-ALIGNED shared-memory address: the entry address is then one
// shift and one three-input LOP3 ((ki << 3) & 0x7f8 | base), and 2^k goes in through a shift + one add
// -- 5 integer-side instructions per exp instead of 7.  Every non-FP64 instruction costs the loop a
// dispatch cycle, so this is worth 4-5 % of the mixture kernel.
template <int G>
__device__ __forceinline__ void exp_neg_half_fast_batch_s(const double (&e)[G], double (&out)[G], unsigned tab_s) {
    double g[G], q[G], tj[G];
    int kk[G];
#pragma unroll
    for (int c = 0; c < G; ++c) {
        const double z = e[c] * kNegHalfLog2e;
        const double t = z + kRintMagic;
        const int ki = max(__double2loint(t), -1021 * kExp2TabSize);
        g[c] = z - (t - kRintMagic);
        const unsigned addr = (((unsigned) ki << 3) & (unsigned) ((kExp2TabSize - 1) * 8)) | tab_s;
        asm("ld.shared.f64 %0, [%1];" : "=d"(tj[c]) : "r"(addr));
        asm("shr.s32 %0, %1, %2;" : "=r"(kk[c]) : "r"(ki), "n"(kExp2TabBits));  // opaque: keeps it one SHF + one add
    }
#pragma unroll
    for (int c = 0; c < G; ++c) q[c] = fma(kExp2C4, g[c], kExp2C3);
#pragma unroll
    for (int c = 0; c < G; ++c) q[c] = fma(q[c], g[c], kExp2C2);
#pragma unroll
    for (int c = 0; c < G; ++c) q[c] = fma(q[c], g[c], kExp2C1);
#pragma unroll
    for (int c = 0; c < G; ++c) {
        const double p = fma(q[c] * g[c], tj[c], tj[c]);
        out[c] = __hiloint2double(__double2hiint(p) + (kk[c] << 20), __double2loint(p));
    }
}
// ---------------------------------------------------------------------------------------------
// Mixture inner loop, second generation: sum_c exp(-e_c / 2) accumulated straight into two partial sums.
// Every instruction counts here: the kernel is bound by dispatch (an FP64 instruction holds the port for two
// cycles, anything else for one; scripts/fp64_probe.py), so the exp is cut to 7 FP64 + 4 other instructions:
//   * a 2048-entry table B3 2^(j/2048) (16 KB of shared memory) leaves |g| <= 2^-12, where the degree-3 Taylor
//     polynomial of 2^g is exact to 3.4e-17 relative;
//   * the table stores each entry with (j << 9) already SUBTRACTED from its high word, so the exponent surgery
//     is a single IMAD: hi' + Z * 512 == hi + ((Z >> 11) << 20) for Z = rint(2048 z) = 2048 k + j;
//   * the result is never formed on its own: lp += (2^k tab_j) * P(g) is one FMA (P = 1 + g(...), 1 ulp);
//   * the argument is clamped on the integer pipe BEFORE the rounding trick (one min on the high word of
//     e >= 0, in place): Z can neither wrap (the 256-entry version wrapped for |theta - x| > 3400) nor leave
//     the normal range.  A clamped term is ~2^-1021: a row that is NaN / inf / absurdly far in EVERY
//     component therefore sums to < 2^-959 and is re-done exactly by the slow path (needs_slow_path()); a NaN
//     in a single component of theta would be lost, so the kernel flags non-finite thetas when it stages them
//     and poisons that node's sums itself (what numpy does: NaN term -> NaN row -> NaN item).
// ---------------------------------------------------------------------------------------------
constexpr int kExpBigBits = 11;
constexpr int kExpBigSize = 1 << kExpBigBits;
constexpr double kRintMagicBig = 3298534883328.0;  // 1.5 * 2^41: rint(2048 z) lands in the low word
constexpr int kExpArgClampHi = 0x40960000;         // high word of 1408.0: -e/2 log2(e) >= -1015.7 (entries are >= 2^-5)
// 2^g = B3 (g^3 + Q2 g^2 + Q1 g + Q0) on |g| <= 2^-12 (truncation 3.4e-17): the MONIC form has exactly one constant
// per instruction (a DADD and two FMAs; an FMA with two constants needs one of them in a register pair, which
// ptxas re-materialises with two MOVs per iteration), and the leading coefficient B3 = ln2^3/6 lives in the table.
constexpr double kExpQ2 = 4.328085122666891, kExpQ1 = 12.488213886033646, kExpQ0 = 18.016684242941434;
// filled once per device by the host (pf_pointwise.cu: ensure_exp_table): long-double B3 2^(j/2048), high word - (j << 9)
static __device__ __align__(16) double g_exp2_big[kExpBigSize];
// ... and the log table as log_pos_fast_k wants it ({1/c_j doubled for j >= kLogTabFirstHalf, log c_j}), so that both
// tables reach shared memory by one TMA bulk copy each instead of ten dependent loads per thread
static __device__ __align__(16) double g_log_adj[512];

template <int G>
__device__ __forceinline__ void mix_exp_accumulate(double (&e)[G], double& lp0, double& lp1, unsigned tab_s) {
    double g[G], tjs[G];
#pragma unroll
    for (int c = 0; c < G; ++c) {
        // clamp in place (one VIMNMX on the high word; e >= 0 or NaN)
        e[c] = __hiloint2double(min(__double2hiint(e[c]), kExpArgClampHi), __double2loint(e[c]));
        const double t = fma(e[c], kNegHalfLog2e, kRintMagicBig);
        const int Z = __double2loint(t);
        g[c] = fma(e[c], kNegHalfLog2e, -(t - kRintMagicBig));
        unsigned idx, addr;
        double tj;
        // LOP3 + IMAD + LDS (left to itself the compiler shifts first: SHL + LOP3 + IADD)
        asm("and.b32 %0, %1, %2;" : "=r"(idx) : "r"(Z), "n"(kExpBigSize - 1));
        asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(addr) : "r"(idx), "r"(tab_s));
        asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(addr));
        tjs[c] = __hiloint2double(__double2hiint(tj) + Z * 512, __double2loint(tj));
    }
    double P[G];
#pragma unroll
    for (int c = 0; c < G; ++c) P[c] = g[c] + kExpQ2;
#pragma unroll
    for (int c = 0; c < G; ++c) P[c] = fma(P[c], g[c], kExpQ1);
#pragma unroll
    for (int c = 0; c < G; ++c) P[c] = fma(P[c], g[c], kExpQ0);
#pragma unroll
    for (int c = 0; c < G; ++c) {
        if (c & 1)
            lp1 = fma(tjs[c], P[c], lp1);
        else
            lp0 = fma(tjs[c], P[c], lp0);
    }
}

// log_pos_fast for the same loop, with as many constants as possible in IMMEDIATE form (a DFMA immediate is the
// high word of a double whose low word is zero): 1/5 and 1/6 are rounded to 21 significant bits (their terms are
// <= 2^-40 r), ln 2 is split as a 21-bit head (n * head exact) + a full-precision tail.  Fewer 64-bit literals in
// the loop means fewer UMOV pairs with which ptxas re-materialises them every iteration.
constexpr double kLn2Head = 0.693147182464599609375;              // 0x3FE62E4300000000
constexpr double kLn2Tail = -1.904654299957767878541823e-09;       // ln 2 - kLn2Head
constexpr double kLogC5 = 0.2000000476837158203125;               // 0x3FC9999A00000000 ~ 1/5
constexpr double kLogC6 = -0.16666662693023681640625;             // 0xBFC5555500000000 ~ -1/6
// The table handed to this version has 1/c_j DOUBLED for j >= kLogTabFirstHalf, because m = x 2^-n is formed with
// the n that already includes the "+1" of the upper half (m in [0.707, 1.414)): adding 0x96000 to the high word
// carries into the exponent exactly when the mantissa is >= 106/256, so n and m cost 4 integer instructions.
__device__ __forceinline__ double log_pos_fast_k(double x, unsigned ltab_s) {
    const int hi = __double2hiint(x);
    const int A = hi + ((0x00100000 - (kLogTabFirstHalf << 12)) - 0x3FF00000);
    const int n = A >> 20;
    const double m = __hiloint2double(hi - (A & (int) 0xFFF00000), __double2loint(x));
    double2 t;  // {1/c_j, log c_j}: the table is addressed by its 32-bit shared-memory address (no pointer to re-form)
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "r"(ltab_s + ((unsigned) (hi >> 8) & 0xFF0u)));
    const double r = fma(m, t.x, -1.0);
    double p = fma(kLogC6, r, kLogC5);
    p = fma(p, r, -0.25);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    const double l1p = fma(r * r, p, r);
    const double nd = (double) n;
    return fma(nd, kLn2Head, fma(nd, kLn2Tail, t.y + l1p));
}

// ---------------------------------------------------------------------------------------------
// Mixture inner loop, third generation (the default).  What bounds the loop is measured, not assumed (ncu, round 2):
// with the 2048-entry table the SHARED-MEMORY pipe carried 0.91 wavefronts per cycle per SM -- 59 % of them bank-conflict
// replays of the eight random 8-byte table reads per (row, node): a half-warp's 16 lanes hit 16 bank pairs at random, so a
// read takes ~6.5 wavefronts instead of 2 -- and removing FP64 instructions alone changed nothing (6.00 -> 5.95 ms).
//
// (a) CONFLICT-FREE table: 256 entries, each stored 16 times side by side (g_exp2_rep[j][c], 32 KB of shared memory), and
//     lane l reads copy c = l mod 16, i.e. always bank pair l mod 16: every lookup is exactly 2 wavefronts.  With 256
//     entries |g| <= 2^-9, where the degree-3 MINIMAX polynomial of 2^g (Chebyshev economisation of the Taylor series)
//     is exact to 1.8e-14 relative.  Monic: 2^g = B3 (g^3 + Q2 g^2 + Q1 g + Q0), B3 folded into the table -- one constant
//     per instruction (an FMA with two constants costs two re-materialising moves); entries carry (j << 12) subtracted
//     from their high word so the exponent surgery is one IMAD (hi' + Z * 4096, Z = rint(256 z) = 256 k + j).
// (b) "Factored" exponent (mix_exp_accumulate_z): |theta_k - x|^2 = |theta_k|^2 - 2 theta_k.x + |x|^2, and the |x|^2 part is
//     the same for every component, so   log sum_k exp(-|theta_k - x|^2 / 2) = -|x|^2/2 + ln sum_k 2^(z_k),
//         z_k = A_k + B_k . x,   A_k = -log2(e)/2 |theta_k|^2,   B_k = log2(e) theta_k   (per node, staged once per call):
//     d FMAs per component instead of d subtractions + d multiply-adds + the scaling.  K = 8, d = 2: 9 FP64 instructions
//     per component (2 z, 3 range reduction, 3 polynomial, 1 accumulate) against 11 before.  Rounding: z carries an absolute
//     error of a few 2^-53 max(|A_k|, |B_k.x|); the kernel takes this path only when  |A_k| + sum_j |B_kj| max|x_j| <= 500
//     for every component of every node (theta against the data set's bounding box, checked while theta is staged), which
//     bounds the error of a term by 2e-13 relative in the worst corner and ~1e-15 for data like the benchmark's, keeps
//     every 2^z a normal number, so no clamp, no slow path and no per-row range test are needed; anything else (far
//     thetas, NaN / inf anywhere) runs the clamped distance form below (mix_exp_accumulate_e).
// (c) log (log_pos_acc): degree-4 series on |r| <= 2^-9 (truncation r^5/5 <= 5.7e-15 absolute), and the exponent n is
//     accumulated as an INTEGER over the lane's rows of a tile; n ln 2 is added once per tile -- 5 FP64 + the two
//     accumulating adds per row instead of 10 + 1.
// ---------------------------------------------------------------------------------------------
constexpr int kExpRepBits = 8, kExpRepSize = 1 << kExpRepBits, kExpRepCopies = 16;
constexpr double kExp3Q2 = 4.3280852879260374, kExp3Q1 = 12.488212455522227, kExp3Q0 = 18.016682179149544;
constexpr long double kExp3B3 = 0.05550411502275755523220807L;  // leading coefficient (in the table)
constexpr double kRintMagicRep = 26388279066624.0;              // 1.5 * 2^44: rint(256 z) lands in the low word
constexpr double kFactoredLimit = 500.0;  // bound on |z| under which the factored form is used (the product of two row
                                          // sums, each in [2^-500, 32 * 2^500], stays a normal number: PF_MIX_PROD)
// filled once per device by the host (ensure_exp_table): [256][16] copies of (long-double B3 2^(j/256), high word - (j << 12))
static __device__ __align__(128) double g_exp2_rep[kExpRepSize * kExpRepCopies];

// one component: t = z + magic already formed; lp += 2^k T[j] (g^3 + Q2 g^2 + Q1 g + Q0).  `tab_l` is the shared-memory
// address of THIS LANE's copy of entry 0 (table base + 8 (lane mod 16)); entries are 128 bytes apart.
#define PF_MIX_EXP_CORE(tc, tjsc)                                                                 \
    {                                                                                             \
        const int Z = __double2loint(tc);                                                         \
        unsigned idx, addr;                                                                       \
        double tj;                                                                                \
        asm("and.b32 %0, %1, %2;" : "=r"(idx) : "r"(Z), "n"(kExpRepSize - 1));                    \
        asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(addr) : "r"(idx), "r"(tab_l));                   \
        asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(addr));                                    \
        tjsc = __hiloint2double(__double2hiint(tj) + Z * 4096, __double2loint(tj));               \
    }
#define PF_MIX_EXP_POLY(G)                                                 \
    double P[G];                                                           \
    _Pragma("unroll") for (int c = 0; c < G; ++c) P[c] = g[c] + kExp3Q2;   \
    _Pragma("unroll") for (int c = 0; c < G; ++c) P[c] = fma(P[c], g[c], kExp3Q1); \
    _Pragma("unroll") for (int c = 0; c < G; ++c) P[c] = fma(P[c], g[c], kExp3Q0);

// z given (|z| <= kFactoredLimit guaranteed by the caller): 7 FP64 + 4 other instructions per component
template <int G, bool TWO_ACC>
__device__ __forceinline__ void mix_exp_accumulate_z(const double (&z)[G], double& lp0, double& lp1, unsigned tab_l) {
    double g[G], tjs[G];
#pragma unroll
    for (int c = 0; c < G; ++c) {
        const double t = z[c] + kRintMagicRep;
        g[c] = z[c] - (t - kRintMagicRep);
        PF_MIX_EXP_CORE(t, tjs[c])
    }
    PF_MIX_EXP_POLY(G)
#pragma unroll
    for (int c = 0; c < G; ++c) {
        if (TWO_ACC && (c & 1))
            lp1 = fma(tjs[c], P[c], lp1);
        else
            lp0 = fma(tjs[c], P[c], lp0);
    }
}

// R rows walking G / R components together: chain c belongs to row c % R
template <int G, int R>
__device__ __forceinline__ void mix_exp_accumulate_zr(const double (&z)[G], double (&lp)[R], unsigned tab_l) {
    double g[G], tjs[G];
#pragma unroll
    for (int c = 0; c < G; ++c) {
        const double t = z[c] + kRintMagicRep;
        g[c] = z[c] - (t - kRintMagicRep);
        PF_MIX_EXP_CORE(t, tjs[c])
    }
    PF_MIX_EXP_POLY(G)
#pragma unroll
    for (int c = 0; c < G; ++c) lp[c % R] = fma(tjs[c], P[c], lp[c % R]);
}

// e = squared distance >= 0 (or NaN), clamped in place like mix_exp_accumulate: 7 FP64 + 5 other per component
template <int G>
__device__ __forceinline__ void mix_exp_accumulate_e(double (&e)[G], double& lp0, double& lp1, unsigned tab_l) {
    double g[G], tjs[G];
#pragma unroll
    for (int c = 0; c < G; ++c) {
        e[c] = __hiloint2double(min(__double2hiint(e[c]), kExpArgClampHi), __double2loint(e[c]));
        const double t = fma(e[c], kNegHalfLog2e, kRintMagicRep);
        g[c] = fma(e[c], kNegHalfLog2e, -(t - kRintMagicRep));
        PF_MIX_EXP_CORE(t, tjs[c])
    }
    PF_MIX_EXP_POLY(G)
#pragma unroll
    for (int c = 0; c < G; ++c) {
        if (c & 1)
            lp1 = fma(tjs[c], P[c], lp1);
        else
            lp0 = fma(tjs[c], P[c], lp0);
    }
}

// log x = n ln 2 + (log c_j + log1p(r)): returns the bracket, adds n to `nacc` (the caller adds nacc ln 2 once per tile:
// log_n_times_ln2).  x must be a positive normal number.  Its table (g_log_mid, filled by the host) has the MIDPOINT
// of every mantissa interval as c_j -- also for the two intervals next to 1, which g_log_adj pins to c = 1 so that
// log(1) = 0 exactly: |r| <= 2^-9 everywhere (with c = 1, r reaches 2^-8 and r^5/5 = 1.8e-13); log(1) comes out as
// ~1e-17 instead of 0, which a sum over rows does not see.  Same addressing and same doubling of 1/c_j for
// j >= kLogTabFirstHalf as log_pos_fast_k.
static __device__ __align__(16) double g_log_mid[512];
constexpr double kLogC3 = 0.33333325386047363281;  // 0x3FD5555500000000 ~ 1/3 (its term is <= 2.5e-9: error 6e-16)
__device__ __forceinline__ double log_pos_acc(double x, unsigned ltab_s, int& nacc) {
    const int hi = __double2hiint(x);
    const int A = hi + ((0x00100000 - (kLogTabFirstHalf << 12)) - 0x3FF00000);
    nacc += A >> 20;
    const double m = __hiloint2double(hi - (A & (int) 0xFFF00000), __double2loint(x));
    double2 t;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "r"(ltab_s + ((unsigned) (hi >> 8) & 0xFF0u)));
    const double r = fma(m, t.x, -1.0);
    double p = fma(-0.25, r, kLogC3);
    p = fma(p, r, -0.5);
    const double l1p = fma(r * r, p, r);
    return t.y + l1p;
}
__device__ __forceinline__ double log_n_times_ln2(int nacc, double rest) {
    const double nd = (double) nacc;
    return fma(nd, kLn2Head, fma(nd, kLn2Tail, rest));  // |nacc| < 2^21: nd * kLn2Head is exact
}

template <int G>
__device__ __forceinline__ void exp_neg_half_fast_batch_s(const float (&e)[G], float (&out)[G], unsigned) {
#pragma unroll
    for (int c = 0; c < G; ++c) out[c] = exp2f(e[c] * -0.72134752f);
}

template <int G>
__device__ __forceinline__ void exp_neg_half_fast_batch(const float (&e)[G], float (&out)[G], const double*) {
#pragma unroll
    for (int c = 0; c < G; ++c) out[c] = exp2f(e[c] * -0.72134752f);
}

__device__ __forceinline__ float exp_neg_half_fast(float e, const double*) {
    return exp2f(e * -0.72134752f);  // MUFU.EX2; underflows to 0 by itself
}

// true when a mixture row sum needs the slow path: it is tiny (< 2^-959: flushed terms could
// matter), zero, negative, inf or nan.
__device__ __forceinline__ bool needs_slow_path(double lp) {
    return (unsigned) (__double2hiint(lp) - 0x04000000) >= 0x7BF00000u;
}
__device__ __forceinline__ bool needs_slow_path(float lp) { return !(lp >= 1e-30f && lp < 3e38f); }
// The same test accumulated over many rows with one VIADD + one VIMNMX.U32 per row: acc = max(acc, key(lp)),
// tested once per tile (acc_needs_slow_path).
__device__ __forceinline__ void slow_path_accumulate(unsigned& acc, double lp) {
    acc = max(acc, (unsigned) (__double2hiint(lp) - 0x04000000));
}
__device__ __forceinline__ void slow_path_accumulate(unsigned& acc, float lp) { acc |= needs_slow_path(lp) ? 0x80000000u : 0u; }
__device__ __forceinline__ bool acc_needs_slow_path(unsigned acc) { return acc >= 0x7BF00000u; }

// log(x) for x > 0 normal (callers route everything else through needs_slow_path / log()).
// x = 2^n m; j = top 8 mantissa bits; r = m / c_j - 1 (one FMA, |r| <= 2^-8); log x = n ln2 + log c_j + log1p(r),
// log1p by its degree-6 series: 10 FP64 instructions, one LDS.128, dependency depth ~9.
// Absolute error <= 5e-16 (relative <= 6e-14 next to x = 1, exact at 1): what a SUM of logs needs.
__device__ __forceinline__ double log_pos_fast(double x, const double2* tab) {
    const int hi = __double2hiint(x);
    const int j = (hi >> 12) & 255;
    const int n = (hi >> 20) - 1023 + (j >= kLogTabFirstHalf ? 1 : 0);
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(x));
    const double2 t = tab[j];
    const double r = fma(m, t.x, -1.0);
    double p = -1.0 / 6.0;
    p = fma(p, r, 0.2);
    p = fma(p, r, -0.25);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    const double l1p = fma(r * r, p, r);
    const double nd = (double) n;
    return fma(nd, kLn2Hi, fma(nd, kLn2Lo, t.y + l1p));  // n ln2 (hi + lo)
}

// series version (no table): x = 2^n m, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m-1)/(m+1),
// odd series through s^21 (<= 2 ulp).  Used by the full-range wrapper below.
__device__ __forceinline__ double log_pos_series(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int n = (hi >> 20) - 1023;
    hi = (hi & 0x000FFFFF) | 0x3FF00000;  // m in [1, 2)
    const bool big = hi >= 0x3FF6A09F;    // m >= ~sqrt(2): halve
    hi = big ? hi - 0x00100000 : hi;
    n = big ? n + 1 : n;
    const double m = __hiloint2double(hi, lo);
    const double num = m - 1.0, den = m + 1.0;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));  // ~2^-23; two Newton steps -> full precision
    double err = fma(-den, r, 1.0);
    r = fma(r, err, r);
    err = fma(-den, r, 1.0);
    r = fma(r, err, r);
    const double s = num * r;
    const double s2 = s * s;
    // q(s2) = sum_{i<10} s2^i / (2i+3), evaluated as A(s4) + s2 B(s4): two independent Horner chains
    const double s4 = s2 * s2;
    double qa = kAtanhC[8], qb = kAtanhC[9];
#pragma unroll
    for (int i = 6; i >= 0; i -= 2) {
        qa = fma(qa, s4, kAtanhC[i]);
        qb = fma(qb, s4, kAtanhC[i + 1]);
    }
    const double q = fma(qb, s2, qa);
    const double s3 = s * s2;
    const double logm = fma(s3 + s3, q, s + s);  // 2s + 2 s^3 (1/3 + s^2/5 + ...)
    const double nd = (double) n;
    return fma(nd, kLn2Hi, fma(nd, kLn2Lo, logm));  // n ln2 (hi + lo)
}

__device__ __forceinline__ float log_pos_fast(float x) {
    return __logf(x);  // MUFU.LG2 * ln2
}
__device__ __forceinline__ float log_pos_fast(float x, const double2*) { return __logf(x); }

// full-range wrappers (diagnostics and slow paths)
__device__ __forceinline__ double exp_neg_half(double e, const double* tab) {
    const double z = e * -0.72134752044448170368;
    if ((unsigned) (__double2hiint(z) & 0x7FFFFFFF) >= 0x408FE800u) return exp(-0.5 * e);
    return exp_neg_half_fast(e, tab);
}
__device__ __forceinline__ float exp_neg_half(float e, const double* tab) { return exp_neg_half_fast(e, tab); }
__device__ __forceinline__ double log_pos(double x) {
    if ((unsigned) (__double2hiint(x) - 0x00100000) >= 0x7FE00000u) return log(x);  // 0, subnormal, inf, nan, < 0
    return log_pos_series(x);
}
__device__ __forceinline__ float log_pos(float x) { return log_pos_fast(x); }
__device__ __forceinline__ double log_pos_tab(double x, const double2* tab) {
    if ((unsigned) (__double2hiint(x) - 0x00100000) >= 0x7FE00000u) return log(x);
    return log_pos_fast(x, tab);
}

}  // namespace pf
